"""A fixed-seed slice of the randomised parity sweep (scripts/fuzz_parity.py) inside the -m gpu gate.

Random scenes — every shape, meshes with / without uvs and opacity maps, depth-shift textures, transparent / reflective /
refractive nodes, point and area lights, random cameras, 1-3 spp — rendered on the device through the C-ABI under random
driver / kernel knobs, against the f64 oracle.

Bar per scene: every over-tolerance pixel is a classified edge flip (util.edge_flip_report) and the ray counts agree to
0.5 %.  A scene that misses that bar must be PROVEN precision-chaotic, pixel by pixel: each unexplained pixel is one where
the oracle's own f64 and f32 modes disagree by more than the tolerance, or where the device agrees with the f32 twin (same
arithmetic, so the difference to f64 is rounding, not logic).  Anything else fails.  Over all scenes the over-tolerance
fraction is printed and must stay below 3e-3 (tiny images: one pixel of a 17x9 frame is 6.5e-3 on its own).
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import fuzz_parity as FZ  # noqa: E402
import oracle_lib as O  # noqa: E402
from nrays_b200 import Scene, make_camera  # noqa: E402
from util import TOL, _nbhd_min_max, edge_flip_report, look, render_both, unexplained_allowed  # noqa: E402

pytestmark = pytest.mark.gpu

COUNT_KEYS = ("rays_primary", "rays_reflect", "rays_refract", "rays_shadow", "paths_truncated")


def _counts_ok(st, ost):
    return all(abs(int(getattr(st, k)) - int(getattr(ost, k))) <= max(6, 5e-3 * int(getattr(ost, k))) for k in COUNT_KEYS)


@pytest.mark.parametrize("seed,n_scenes", [(1, 40), (7, 40)])
def test_fuzz_slice(gpu, seed, n_scenes):
    rng = np.random.default_rng(seed)
    tot_px = tot_over = chaotic = 0
    for i in range(n_scenes):
        nodes, lights = FZ.rand_scene(rng)
        eye = tuple(rng.uniform(-1, 1, 3) * 2.0 + np.array([0.0, 1.0, -7.0]))
        w, h = int(rng.integers(17, 80)), int(rng.integers(9, 60))
        spp = int(rng.integers(1, 4))
        window = 0.0 if rng.uniform() < 0.4 else 1.0
        knobs = FZ.KNOBS[int(rng.integers(0, len(FZ.KNOBS)))]
        os.environ.update(knobs)
        try:
            img, st, ref, ost = render_both(nodes, lights, eye=eye, w=w, h=h, spp=spp, window=window, seed=i, max_depth=12)
        finally:
            for k in knobs:
                os.environ.pop(k, None)
        assert np.isfinite(img).all(), (seed, i)
        rep = edge_flip_report(img, ref, (w, h))
        tot_px += w * h
        tot_over += rep["over"]
        if rep["unexplained"] <= unexplained_allowed(w * h) and _counts_ok(st, ost):
            continue
        # ---- prove it: precision-chaotic, or fail ------------------------------------------------------------
        sc = Scene(nodes, lights, (1.0, 1.0, 1.0), upload=False)
        cam = make_camera(w, h, spp, window, eye, look(eye, (0, 0, 0), 45.0, w, h), seed=i, max_depth=12)
        ref32, ost32 = O.OracleScene(sc.flat, 32).render(cam)
        a = np.asarray(img, np.float64).reshape(h, w, 3)
        b = np.asarray(ref, np.float64).reshape(h, w, 3)
        c = np.asarray(ref32, np.float64).reshape(h, w, 3)
        lo, hi = _nbhd_min_max(b)
        over = np.abs(a - b).max(axis=2) > TOL
        flip = over & ((((hi - lo).max(axis=2) > 2 * TOL) & ((a >= lo - TOL) & (a <= hi + TOL)).all(axis=2)) | ((hi - lo).max(axis=2) > 0.1))
        bad = over & ~flip
        oracles_disagree = np.abs(b - c).max(axis=2) > TOL
        device_is_twin = np.abs(a - c).max(axis=2) <= TOL
        unproven = bad & ~(oracles_disagree | device_is_twin)
        counts_twin = _counts_ok(st, ost32) or not _counts_ok(ost32, ost)   # counts: agree with the twin, or the oracles disagree
        print("fuzz seed %d scene %d %dx%dx%d knobs=%s: over=%d unexplained=%d -> unproven=%d, f64-vs-f32 oracle over=%d, counts dev/f64/f32 %s" % (
            seed, i, w, h, spp, ",".join(knobs) or "-", rep["over"], rep["unexplained"], int(unproven.sum()), int(oracles_disagree.sum()),
            " ".join("%d/%d/%d" % (getattr(st, k), getattr(ost, k), getattr(ost32, k)) for k in COUNT_KEYS)))
        assert not unproven.any(), "seed %d scene %d: %d pixels off against BOTH oracle modes where the modes agree" % (seed, i, int(unproven.sum()))
        assert counts_twin, "seed %d scene %d: ray counts differ from both oracle modes" % (seed, i)
        chaotic += 1
    frac = tot_over / float(tot_px)
    print("parity: fuzz seed %d: %d scenes, %d pixels, over-tolerance fraction %.5f, %d scenes needed the chaos proof" % (
        seed, n_scenes, tot_px, frac, chaotic))
    assert frac <= 3e-3
    assert chaotic <= n_scenes // 8


def _rand_mesh_scene(rng):
    """Mesh-only random scene (the grid node formats 3 / 4 exist for those): jittered quad grids with random transforms, opacity
    maps, transparency, reflection and refraction, one or two lights."""
    nodes = []
    for _ in range(int(rng.integers(1, 6))):
        P, F, UV = FZ.quad_mesh(float(rng.uniform(0.4, 2.5)), int(rng.integers(1, 14)), y=0.0)
        P = P + rng.normal(0.0, 0.04, P.shape).astype(np.float32)
        if rng.uniform() < 0.3:
            P[:, 1] += 0.4 * np.sin(3.0 * P[:, 0]) * np.cos(2.0 * P[:, 2])   # curved sheet
        alpha = 1.0 if rng.uniform() < 0.6 else float(rng.uniform(0.1, 0.9))
        refl = (0.0, 0.0) if rng.uniform() < 0.6 else (float(rng.uniform(0.1, 0.9)), float(rng.choice([0.2, 0.35, 0.5])))
        refr = 1.0 if rng.uniform() < 0.5 else float(rng.uniform(1.05, 1.8))
        mat = FZ.rand_material(rng, True)
        nodes.append(FZ.node(FZ.TriMesh(P, F, UV), mat, pos=tuple(rng.uniform(-2.0, 2.0, 3)), angle=tuple(rng.uniform(-180, 180, 3)),
                             refl=refl, alpha=alpha, refr=refr))
    lights = []
    for _ in range(int(rng.integers(1, 3))):
        radius = 0.0 if rng.uniform() < 0.6 else float(rng.uniform(0.05, 0.5))
        lights.append(FZ.Light(tuple(rng.uniform(-5, 5, 3) + np.array([0, 5, 0])), radius, int(rng.choice([1, 4, 9])) if radius else 1,
                               tuple(rng.uniform(0.3, 1.0, 3))))
    return nodes, lights


def test_fuzz_mesh_scenes_across_node_formats(gpu):
    """30 random mesh-only scenes: node formats 2 / 3 / 4 (bf16 half extents, 16-bit grid records, speculative loop with predicated
    push / pop) give the frame and the ray counts of format 0 — the boxes of every format only ever grow, so the hits are the same
    triangles — and format 0 meets the oracle bar like every other fuzz scene."""
    rng = np.random.default_rng(11)
    tot_px = tot_over = 0
    honoured = {2: 0, 3: 0, 4: 0}
    for i in range(30):
        nodes, lights = _rand_mesh_scene(rng)
        eye = tuple(rng.uniform(-1, 1, 3) * 2.0 + np.array([0.0, 1.0, -7.0]))
        w, h = int(rng.integers(24, 90)), int(rng.integers(16, 64))
        spp = int(rng.integers(1, 3))
        base = None
        for fmt in (0, 2, 3, 4):
            os.environ["NRB_NODE_FORMAT"] = str(fmt)
            try:
                img, st, ref, ost = render_both(nodes, lights, eye=eye, w=w, h=h, spp=spp, window=1.0, seed=i, max_depth=10)
                used = Scene(nodes, lights, (1.0, 1.0, 1.0)).build_info().node_format
            finally:
                os.environ.pop("NRB_NODE_FORMAT", None)
            assert used in (fmt, 0), (i, fmt, used)   # 0: a mesh without inner nodes, or a grid that cannot resolve the scene
            assert np.isfinite(img).all(), (i, fmt)
            if fmt and used == fmt:
                honoured[fmt] += 1
            if base is None:
                base = (img, st)
                rep = edge_flip_report(img, ref, (w, h))
                tot_px += w * h
                tot_over += rep["over"]
                continue
            d = np.abs(base[0] - img).max(axis=1)
            assert (d > 1e-4).mean() < 5e-3, (i, fmt, float((d > 1e-4).mean()))   # identical hits up to exact ties
            assert abs(int(base[1].rays_total) - int(st.rays_total)) <= max(4, 2e-3 * base[1].rays_total), (i, fmt)
    frac = tot_over / float(tot_px)
    print("parity: mesh fuzz: 30 scenes x 4 node formats, %d pixels, over-tolerance fraction of format 0 vs the oracle %.5f" % (tot_px, frac))
    assert frac <= 3e-3
    assert min(honoured.values()) >= 25, honoured   # the forced format was really the one that rendered
