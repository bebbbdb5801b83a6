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
    scene, cam, cfg = configs.build(name, **kw)
    for env in ({}, {"NRB_TAIL_RAYS": "0"}, {"NRB_BATCH_SLOTS": "1024", "NRB_SHADOW_CAP": "512"},
                {"NRB_REFILL_PRIMARY": "20", "NRB_REFILL_RAYS": "24", "NRB_REFILL_SHADOW": "24"}, {"NRB_REVERSE_SHADOW": "0"}):
        os.environ.update(env)
        img, st = render_camera(scene, cam, resolution=res, aa=aa, seed=1, return_stats=True)
        for k in env:
            os.environ.pop(k)
        print(name, env, st.rays_total, float(img.pixels.mean()))
    scene.close()
print("sanitize: done")
