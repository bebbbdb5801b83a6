"""Experiment driver: render frames of a BASELINE config and print the library's per-frame stats."""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nrays_b200 import _abi as A, _lib, configs, make_camera  # noqa: E402

cfgname = sys.argv[1] if len(sys.argv) > 1 else "C3"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 6
lib = _lib.load()
t0 = time.time()
scene, camd, cfg = configs.build(cfgname)
print("scene build+upload %.2fs" % (time.time() - t0))
w, h, spp = cfg["width"], cfg["height"], cfg["spp"]
import torch  # noqa: E402

out = torch.empty(w * h * 3, dtype=torch.float32, device="cuda")
for f in range(frames):
    cam = make_camera(w, h, spp, cfg["window"], camd.eye, camd.projection((w, h)), seed=f)
    st = A.NrbStats()
    if f == frames - 1:
        os.environ["NRB_DUMP_WAVES"] = "1"
    t0 = time.perf_counter()
    _lib.check(lib.nrb_render_device(scene.handle, C.byref(cam), C.c_void_p(out.data_ptr()), C.byref(st)))
    wall = (time.perf_counter() - t0) * 1e3
    d = st.as_dict()
    print("frame %d wall %.3f ms device %.3f trace %.3f other %.3f rays %d -> %.0f Mrays/s launches %d waves %d" % (
        f, wall, d["ms_device"], d["ms_trace"], d["ms_shade"], d["rays_total"], d["rays_total"] / d["ms_device"] / 1e3,
        d["kernel_launches"], d["waves"]))
