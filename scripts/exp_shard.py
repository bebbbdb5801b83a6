"""Experiment: per-wave times of ONE rank's share of a tile-sharded frame (rank 0 of `world`), on one GPU."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from nrays_b200 import configs, dist, make_camera  # noqa: E402

cfgname = sys.argv[1] if len(sys.argv) > 1 else "C3"
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
scene, camd, cfg = configs.build(cfgname)
w, h, spp = cfg["width"], cfg["height"], cfg["spp"]
tpr = dist.tiles_per_rank(w, h, world)
packed = torch.zeros((tpr, 16, 16, 3), dtype=torch.float32, device="cuda")
for f in range(6):
    cam = make_camera(w, h, spp, cfg["window"], camd.eye, camd.projection((w, h)), seed=f)
    if f == 5:
        os.environ["NRB_DUMP_WAVES"] = "1"
    st, n = dist.render_tiles_device(scene, cam, 0, world, packed)
    d = st.as_dict()
    print("frame %d device %.3f ms trace %.3f rays %d launches %d" % (f, d["ms_device"], d["ms_trace"], d["rays_total"], d["kernel_launches"]))
