import time, sys, ctypes as C
sys.path.insert(0,'/root/repo')
from nrays_b200 import configs, _lib
for c in ("C3","C4"):
    t=time.time(); scene, cam, cfg = configs.build_flat(c); t1=time.time()-t
    lib=_lib.load(); h=C.c_void_p()
    t=time.time(); rc=lib.nrb_scene_create(C.byref(scene.flat.desc), 0, C.byref(h)); t2=time.time()-t
    print(c, 'python flatten %.2fs'%t1, 'nrb_scene_create %.2fs rc=%d'%(t2,rc))
