"""Developer tool: build a BASELINE config's BVH on the host (no GPU) and dump it for scripts/bvh_sim.cpp.
usage: NRB_BVH_CNODE=.. python scripts/bvh_dump.py C3 /tmp/c3.bin"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nrays_b200 import _abi as A, _lib, configs  # noqa: E402

cfg, out = sys.argv[1], sys.argv[2]
scene, cam, c = configs.build_flat(cfg)
os.environ["NRB_DUMP_BVH"] = out
info = A.NrbBuildInfo()
t0 = time.time()
_lib.check(_lib.load().nrb_scene_validate(C.byref(scene.flat.desc), C.byref(info)))
print("nodes %d depth %d build %.0f ms (%.2fs wall)" % (info.bvh_nodes, info.max_depth, info.build_ms, time.time() - t0))
