"""Small renders of every kernel variant for compute-sanitizer (memcheck / racecheck)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from nrays_b200 import configs  # noqa: E402
from nrays_b200.loader3d import render_camera  # noqa: E402

for name, kw, res, aa in (("C1", dict(globe_size=(32, 16)), (48, 40), (2, 1.0)), ("C2", dict(globe_size=(32, 16)), (40, 40), (1, 0.0)),
                          ("C3", dict(target_tris=12000, lod=8), (64, 36), (2, 1.0)), ("C4", dict(target_tris=16000), (48, 28), (1, 0.0))):
  for fmt, builder in (("0", "sah"), ("2", "sah"), ("0", "ploc"), ("2", "lbvh"), ("3", "sah"), ("4", "ploc")):
    if fmt in ("3", "4") and name in ("C1", "C2"):
        continue                             # grid formats: mesh-only scenes
    os.environ["NRB_NODE_FORMAT"] = fmt      # 2 = bf16 half extents + speculative loop; 3 / 4 = 16-bit grid records, plain / speculative loop
    os.environ["NRB_BUILDER"] = builder      # device builders (LBVH / PLOC) run their own kernels
    scene, cam, cfg = configs.build(name, **kw)
    for env in ({}, {"NRB_TAIL_RAYS": "0"}, {"NRB_BATCH_SLOTS": "1024", "NRB_SHADOW_CAP": "512"},
                {"NRB_REFILL_PRIMARY": "20", "NRB_REFILL_RAYS": "24", "NRB_REFILL_SHADOW": "24"}, {"NRB_REVERSE_SHADOW": "0"}):
        os.environ.update(env)
        img, st = render_camera(scene, cam, resolution=res, aa=aa, seed=1, return_stats=True)
        for k in env:
            os.environ.pop(k)
        print(name, fmt, builder, env, st.rays_total, float(img.pixels.mean()))
    scene.close()
# tile exchanges (float and RGB8, device image and host segments)
import ctypes as C  # noqa: E402

import numpy as np  # noqa: E402
import torch  # noqa: E402

from nrays_b200 import dist, make_camera  # noqa: E402

os.environ.pop("NRB_NODE_FORMAT", None)
os.environ.pop("NRB_BUILDER", None)
scene, camd, cfg = configs.build("C3", target_tris=12000, lod=8)
w, h = 64, 48
cam = make_camera(w, h, 2, 1.0, camd.eye, camd.projection((w, h)), seed=3)
img = torch.zeros(w * h * 3, dtype=torch.float32, device="cuda")
img8 = torch.zeros(w * h * 3, dtype=torch.uint8, device="cuda")
for r in range(4):
    dist.render_tiles_to_image(scene, cam, r, 4, img.data_ptr())
    dist.render_tiles_to_image_rgb8(scene, cam, r, 4, img8.data_ptr())
print("tiles", float(img.mean()), float(img8.float().mean()))
scene.close()
print("sanitize: done")
