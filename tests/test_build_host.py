"""Host half of Scene::new through the C-ABI (nrb_scene_validate): table validation and the BVH builder's
structural invariants (checked inside the library: every triangle in exactly one leaf, child boxes inside
their parent's, every node reachable once, and the device copy of every box — centre + half extent, rounded
up — contains the builder's box).  Runs without a GPU."""
import ctypes as C
import os

import numpy as np
import pytest

from nrays_b200 import Ball, Isometry3, NormalMaterial, Plane, Scene, SceneNode, TriMesh, _abi as A, _lib, configs
from util import node, quad_mesh


def validate(flat):
    info = A.NrbBuildInfo()
    rc = _lib.load().nrb_scene_validate(C.byref(flat.desc), C.byref(info))
    return rc, info, _lib.load().nrb_last_error().decode()


def test_validate_baseline_scenes():
    for name, kw, tris, cands in (("C1", dict(globe_size=(32, 16)), 0, 3), ("C2", dict(globe_size=(32, 16)), 0, 0),
                                  ("C3", dict(target_tris=30000, lod=4), 30000, 2), ("C4", dict(target_tris=80000), 80000, 0)):
        scene, _cam, _cfg = configs.build_flat(name, **kw)
        rc, info, err = validate(scene.flat)
        assert rc == A.NRB_OK, err
        assert info.triangles == tris and info.transparent_candidates == cands
        assert info.max_depth <= 60
        if tris:
            assert tris / 4 <= info.bvh_nodes <= tris   # leaves hold <= 4 triangles


def test_parallel_build_is_deterministic_and_valid():
    scene, _cam, _cfg = configs.build_flat("C4", target_tris=320000)
    counts = []
    for th in ("1", "3", "8"):
        os.environ["NRB_BVH_THREADS"] = th
        rc, info, err = validate(scene.flat)
        assert rc == A.NRB_OK, err
        counts.append((info.bvh_nodes, info.max_depth))
    os.environ.pop("NRB_BVH_THREADS")
    assert len(set(counts)) == 1, counts


def test_degenerate_meshes_build():
    P, F, UV = quad_mesh(1.0, 3)
    dup = TriMesh(np.concatenate([P] * 4), np.concatenate([F + k * len(P) for k in range(4)]), np.concatenate([UV] * 4))  # coincident triangles
    sliver = TriMesh(np.array([[0, 0, 0], [1, 0, 0], [2, 0, 0], [0, 1e-30, 0]], np.float32), np.array([[0, 1, 2], [0, 1, 3]], np.uint32))
    rc, info, err = validate(Scene([node(dup, NormalMaterial()), node(sliver, NormalMaterial(), alpha=0.5)], [], upload=False).flat)
    assert rc == A.NRB_OK, err
    assert info.triangles == 18 * 4 + 2 and info.transparent_candidates == 1


def test_malformed_tables_are_rejected_without_a_device():
    def flat():
        P, F, UV = quad_mesh(1.0, 1)
        return Scene([node(Ball(1.0), NormalMaterial()), node(TriMesh(P, F, UV), NormalMaterial())], [], upload=False).flat

    f = flat(); f.node_rows[0].material = 5
    assert validate(f)[0] == A.NRB_ERR_INVALID_ARG
    f = flat(); f.node_rows[0].shape = 42
    assert validate(f)[0] == A.NRB_ERR_INVALID_ARG
    f = flat(); f.node_rows[1].tri_count = 100
    assert validate(f)[0] == A.NRB_ERR_INVALID_ARG
    f = flat(); f.indices[2] = 77
    assert validate(f)[0] == A.NRB_ERR_INVALID_ARG
    f = flat(); f.node_rows[0].nmap_texture = 5     # no such texture
    assert validate(f)[0] == A.NRB_ERR_INVALID_ARG
    f = flat(); f.desc.struct_size = 8
    assert validate(f)[0] == A.NRB_ERR_INVALID_ARG
    f = flat(); f.mat_rows[0].kind = 9
    assert validate(f)[0] == A.NRB_ERR_UNSUPPORTED
    assert _lib.load().nrb_scene_validate(None, None) == A.NRB_ERR_INVALID_ARG
    rc, info, err = validate(flat())
    assert rc == A.NRB_OK and info.shapes == 1 and info.triangles == 2


def test_planes_stay_outside_the_bvh():
    nodes = [node(Plane((0, 1, 0)), NormalMaterial(), pos=(0, -1, 0)), node(Plane((1, 0, 0)), NormalMaterial(), alpha=0.5),
             node(Ball(1.0), NormalMaterial())]
    rc, info, err = validate(Scene(nodes, [], upload=False).flat)
    assert rc == A.NRB_OK and info.planes == 2 and info.shapes == 3 and info.transparent_candidates == 0 and info.bvh_nodes == 0


def test_fast_division_by_frame_constants_is_exact(tmp_path):
    """The kernels divide by spp / tiles_x / width with a precomputed 64-bit reciprocal (device_types.cuh: FastDiv).
    The same struct and formula, compiled for the host, must equal `/` for every divisor the frame can produce."""
    import shutil
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cuda_inc = "/usr/local/cuda/include"
    if not (shutil.which("g++") and os.path.exists(os.path.join(cuda_inc, "cuda_runtime.h"))):
        pytest.skip("needs g++ and the CUDA headers")
    src = tmp_path / "fd.cpp"
    src.write_text(r'''
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include "device_types.cuh"
using namespace nrb;
static uint32_t fdiv(uint32_t n, const FastDiv &f) {  // == device_math.cuh: fdiv
  if (f.d == 1u) return n;
  unsigned long long lo = (unsigned long long)f.mul_lo * n;
  unsigned long long hi = (unsigned long long)f.mul_hi * n + (lo >> 32);
  return (uint32_t)(hi >> 32);
}
int main() {
  unsigned long long bad = 0;
  uint32_t ds[] = {1, 2, 3, 4, 5, 7, 8, 15, 16, 17, 120, 121, 240, 255, 256, 1000, 1080, 1920, 3840, 65535, 65536, 1u << 31, 0x7fffffffu, 0xffffffffu};
  for (uint32_t d : ds) {
    FastDiv f = make_fastdiv(d);
    for (unsigned long long n = 0; n < (1ull << 32); n += 65521) bad += fdiv((uint32_t)n, f) != (uint32_t)n / d;
    uint32_t edge[] = {0u, 1u, d - 1, d, d + 1, 2 * d - 1, 2 * d, 0xfffffffeu, 0xffffffffu, 0xffffffffu - d, 0xffffffffu / d * d, 0xffffffffu / d * d - 1};
    for (uint32_t n : edge) bad += fdiv(n, f) != n / d;
  }
  for (uint32_t d = 1; d < 5000; ++d) {
    FastDiv f = make_fastdiv(d);
    for (uint32_t n = 0; n < 300000; n += 7) bad += fdiv(n, f) != n / d;
    bad += fdiv(0xffffffffu, f) != 0xffffffffu / d;
  }
  printf("%llu\n", bad);
  return bad != 0;
}
''')
    exe = tmp_path / "fd"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", cuda_inc, "-I", os.path.join(root, "nrays_b200", "csrc"), "-o", str(exe), str(src)])
    assert subprocess.check_output([str(exe)]).strip() == b"0"
