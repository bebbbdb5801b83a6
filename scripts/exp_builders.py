"""SAH (host) vs LBVH / PLOC (device) builders: build time and frame time on a BASELINE config.
NRB_PLOC_RADIUS selects the PLOC search radius (default 16); NRB_BUILD_TIMES=1 prints the host phases of Scene::new."""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nrays_b200 import Scene, _abi as A, _lib, configs, make_camera  # noqa: E402
from nrays_b200.loader3d import parse  # noqa: E402

import torch  # noqa: E402

for cfgname in sys.argv[1:] or ["C3", "C4"]:
    cfg = configs.CONFIGS[cfgname]
    lights, nodes, cams = parse(cfg["text"](), cfg["resolver"]())
    w, h, spp = cfg["width"], cfg["height"], cfg["spp"]
    out = torch.empty(w * h * 3, dtype=torch.float32, device="cuda")
    for builder in (os.environ.get("EXP_BUILDERS", "sah,lbvh,ploc").split(",")):
        t0 = time.time()
        scene = Scene(nodes, lights, (1, 1, 1), builder=builder)
        wall = time.time() - t0
        bi = scene.build_info()
        ms = []
        for f in range(5):
            cam = make_camera(w, h, spp, cfg["window"], cams[0].eye, cams[0].projection((w, h)), seed=f)
            st = A.NrbStats()
            _lib.check(_lib.load().nrb_render_device(scene.handle, C.byref(cam), C.c_void_p(out.data_ptr()), C.byref(st)))
            ms.append(st.ms_device)
        print("%s %-4s create %.2f s (flatten+build %.0f ms, device build kernels %.2f ms) nodes %d depth %d | frame %.3f ms -> %.0f Mrays/s" % (
            cfgname, builder, wall, bi.build_ms, bi.gpu_build_ms, bi.bvh_nodes, bi.max_depth, min(ms[1:]), st.rays_total / min(ms[1:]) / 1e3))
        scene.close()
