"""Cross-check of the CPU oracle (oracle/nrays_oracle.cpp) against the independent brute-force checker
(tests/bruteforce.py: no BVT, no shared code, reads the host Scene objects instead of the flattened tables).

The reference holds nothing that could pin the oracle, so this is the pin it gets: both restate
src/scene.rs:147-339 + the ncollide3d conventions of SURVEY App. B, one through the reference's data structures
(two-level median-split BVT, best-first search), the other by exhaustive search.  Bar: toi / normal / uv agree to
1e-9 relative (f64 both sides), colours to 2e-5 (f32 accumulation order), hit / miss and occluded / lit decisions
agree exactly except where the brute force reports a near-tie (|gap| below 1e-9: order-dependent in the reference
itself, B.2 "first found wins").
"""
import math

import numpy as np
import pytest

import bruteforce as BF
import oracle_lib as O
from nrays_b200 import (Ball, Capsule, Cone, Cuboid, Cylinder, ImageData, Interpolation, Isometry3, Light, NormalMaterial,
                        Overflow, PhongMaterial, Plane, Scene, SceneNode, Texture2d, TriMesh, UVMaterial, camera_projection,
                        make_camera)


def _rand_texture(rng, w, h, interp, overflow, binary_alpha=False):
    px = rng.uniform(0.0, 1.0, (h * w, 4)).astype(np.float32)
    if binary_alpha:
        px[:, 3] = (rng.uniform(0, 1, h * w) > 0.5).astype(np.float32)
    return Texture2d(ImageData(px, (w, h)), interp, overflow)


def _rand_mesh(rng, n_tri, spread=1.5, with_uv=True):
    """A small triangle soup (independent triangles: no shared edges, so ties only where we build them)."""
    c = rng.uniform(-spread, spread, (n_tri, 1, 3))
    P = (c + rng.uniform(-0.7, 0.7, (n_tri, 3, 3))).reshape(-1, 3).astype(np.float32)
    Fc = np.arange(3 * n_tri, dtype=np.uint32).reshape(-1, 3)
    UV = rng.uniform(-0.5, 1.5, (3 * n_tri, 2)).astype(np.float32) if with_uv else None
    return TriMesh(P, Fc, UV)


def _rand_material(rng, textured_ok):
    r = rng.integers(0, 6)
    if r == 0:
        return NormalMaterial()
    if r == 1:
        return UVMaterial()
    tex = alpha = None
    if textured_ok and rng.random() < 0.6:
        tex = _rand_texture(rng, int(rng.integers(2, 9)), int(rng.integers(2, 9)),
                            [Interpolation.Bilinear, Interpolation.Nearest][rng.integers(0, 2)],
                            [Overflow.Wrap, Overflow.ClampToEdges][rng.integers(0, 2)])
    if textured_ok and rng.random() < 0.25:
        alpha = _rand_texture(rng, int(rng.integers(2, 9)), int(rng.integers(2, 9)),
                              [Interpolation.Bilinear, Interpolation.Nearest][rng.integers(0, 2)],
                              [Overflow.Wrap, Overflow.ClampToEdges][rng.integers(0, 2)], binary_alpha=rng.random() < 0.5)
    return PhongMaterial(rng.uniform(0, 0.4, 3), rng.uniform(0.2, 1, 3), rng.uniform(0, 1, 3), tex, alpha,
                         float(rng.uniform(2, 80)))


def _rand_scene(seed, n_nodes=None, allow_solid=False):
    rng = np.random.default_rng(seed)
    n_nodes = n_nodes or int(rng.integers(2, 9))
    nodes = []
    for _ in range(n_nodes):
        k = int(rng.integers(0, 8))
        if k == 0:
            g = Ball(float(rng.uniform(0.3, 1.2)))
        elif k == 1:
            g = Cuboid(rng.uniform(0.2, 1.0, 3))
        elif k == 2:
            g = Cylinder(float(rng.uniform(0.3, 1.0)), float(rng.uniform(0.2, 0.9)))
        elif k == 3:
            g = Cone(float(rng.uniform(0.3, 1.0)), float(rng.uniform(0.2, 0.9)))
        elif k == 4:
            g = Capsule(float(rng.uniform(0.3, 1.0)), float(rng.uniform(0.2, 0.7)))
        elif k == 5:
            g = Plane(rng.normal(size=3))
        else:
            g = _rand_mesh(rng, int(rng.integers(1, 24)), with_uv=rng.random() < 0.8)
        mat = _rand_material(rng, textured_ok=True)
        pos = rng.uniform(-3, 3, 3) if not isinstance(g, Plane) else rng.uniform(-4, -2.5, 3) * np.array([0, 1, 0])
        ang = rng.uniform(-math.pi, math.pi, 3) if rng.random() < 0.7 else np.zeros(3)
        if isinstance(g, Plane):
            ang = np.zeros(3)
        refl = (float(rng.uniform(0.1, 0.6)), float(rng.choice([0.2, 0.35, 0.5]))) if rng.random() < 0.3 else (0.0, 0.0)
        alpha = float(rng.choice([1.0, 1.0, 1.0, 1.0, 0.2, 0.6, 0.0]))
        refr = float(rng.choice([1.0, 1.3, 1.5]))
        solid = bool(allow_solid and rng.random() < 0.3)
        nodes.append(SceneNode(mat, refl[0], refl[1], alpha, refr, Isometry3.new(pos, ang), g, None, solid))
    lights = [Light(rng.uniform(-6, 6, 3) + np.array([0, 6, 0]), 0.0, int(rng.choice([1, 4, 10])), rng.uniform(0.3, 1.0, 3))
              for _ in range(int(rng.integers(0, 3)))]
    return nodes, lights, rng


def _rays(rng, n, aim_spread=3.0):
    o = rng.uniform(-7, 7, (n, 3))
    tgt = rng.uniform(-aim_spread, aim_spread, (n, 3))
    d = tgt - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return o, d


@pytest.mark.parametrize("seed", range(12))
def test_cast_matches_brute_force_per_node(seed):
    """SceneNode::cast for every node kind: toi, normal, uv presence and value (src/scene_node.rs:51-58)."""
    nodes, lights, rng = _rand_scene(1000 + seed, allow_solid=True)
    sc = Scene(nodes, lights, upload=False)
    orc = O.OracleScene(sc.flat, 64)
    bs = BF.BruteScene(nodes, lights)
    o, d = _rays(rng, 160)
    # some origins INSIDE shapes: inside-hit conventions (normal direction, solid flag)
    for k in range(0, len(o), 4):
        n = nodes[(k // 4) % len(nodes)]
        if not isinstance(n.geometry, (Plane, TriMesh)):
            o[k] = n.transform.trans + rng.uniform(-0.15, 0.15, 3)
    checked = hits = 0
    for i in range(len(nodes)):
        # half of the rays are aimed at this node so that most of them hit it
        aim = nodes[i].transform.trans[None, :] + rng.uniform(-0.8, 0.8, (len(o), 3))
        di = aim - o
        di /= np.linalg.norm(di, axis=1, keepdims=True)
        for k in range(len(o)):
            dk = di[k] if (k & 1) else d[k]
            a = orc.cast(i, o[k], dk)
            b = bs.cast(i, o[k], dk)
            checked += 1
            if a is None or b is None:
                assert a is None and b is None, "node %d (%s) ray %d: oracle %r vs brute %r" % (
                    i, type(nodes[i].geometry).__name__, k, a, None if b is None else b.toi)
                continue
            hits += 1
            assert a["toi"] == pytest.approx(b.toi, rel=1e-9, abs=1e-9), (i, type(nodes[i].geometry).__name__, k)
            if b.toi == 0.0 and nodes[i].solid:
                continue  # solid inside hit: toi 0, normal unspecified
            assert np.allclose(a["normal"], b.normal, atol=1e-7), (i, type(nodes[i].geometry).__name__, k, a["normal"], b.normal)
            assert (a["uv"] is None) == (b.uv is None)
            if b.uv is not None:
                assert np.allclose(a["uv"], b.uv, atol=1e-8), (i, type(nodes[i].geometry).__name__, k, a["uv"], b.uv)
    assert hits > 50, "too few hits to mean anything (%d of %d)" % (hits, checked)


@pytest.mark.parametrize("seed", range(10))
def test_intersects_ray_matches_brute_force(seed):
    """Scene::intersects_ray: opaque occlusion vs transparent filter product, per NODE closest hit
    (src/scene.rs:147-161, 304-339)."""
    nodes, lights, rng = _rand_scene(2000 + seed, n_nodes=10)
    sc = Scene(nodes, lights, upload=False)
    orc = O.OracleScene(sc.flat, 64)
    bs = BF.BruteScene(nodes, lights)
    o, d = _rays(rng, 240)
    for k in range(0, len(o), 2):  # every other ray is aimed at some node
        tgt = nodes[(k // 2) % len(nodes)].transform.trans + rng.uniform(-0.5, 0.5, 3)
        d[k] = (tgt - o[k]) / np.linalg.norm(tgt - o[k])
    n_some = n_none = n_filtered = 0
    for k in range(len(o)):
        maxtoi = float(rng.uniform(1.0, 14.0))
        a = orc.intersects_ray(o[k], d[k], maxtoi)
        b, margin = bs.intersects_ray(o[k], d[k], maxtoi)
        if margin < 1e-9:
            continue  # a hit exactly at maxtoi: <= decided by the last bit
        assert (a is None) == (b is None), "ray %d: oracle %r vs brute %r" % (k, a, b)
        if b is None:
            n_none += 1
        else:
            n_some += 1
            n_filtered += int(not np.all(b == 1.0))
            assert np.allclose(a, b, rtol=2e-6, atol=1e-7), (k, a, b)
    assert n_some > 10 and n_none > 10 and n_filtered > 3, (n_some, n_none, n_filtered)


def test_per_node_transparent_shadow_semantics_mesh():
    """A mesh whose FIRST triangle along the ray is transparent (opacity map alpha 0) and whose second is opaque: the
    node's CLOSEST hit decides, so the node filters and does not occlude; two such nodes multiply; an opaque node
    behind still occludes (src/scene.rs:313-337)."""
    # alpha texture: left half (u < 0.5) alpha 0, right half alpha 1 (Nearest, Clamp)
    px = np.ones((2, 4), np.float32)
    px[0, 3] = 0.0
    amap = Texture2d(ImageData(px, (2, 1)), Interpolation.Nearest, Overflow.ClampToEdges)
    mat = PhongMaterial((0.5, 0.25, 0.125), (1, 1, 1), (1, 1, 1), None, amap, 10.0)

    def two_layer(z0):
        # two parallel triangles at z0 and z0 + 1; uvs put the first in the alpha-0 half, the second in the alpha-1 half
        P = np.array([[-2, -2, z0], [2, -2, z0], [0, 2, z0], [-2, -2, z0 + 1], [2, -2, z0 + 1], [0, 2, z0 + 1]], np.float32)
        Fc = np.array([[0, 1, 2], [3, 4, 5]], np.uint32)
        UV = np.array([[0.1, 0.5]] * 3 + [[0.9, 0.5]] * 3, np.float32)
        return TriMesh(P, Fc, UV)

    n1 = SceneNode(mat, 0, 0, 1.0, 1.0, Isometry3.identity(), two_layer(0.0))
    n2 = SceneNode(mat, 0, 0, 1.0, 1.0, Isometry3.identity(), two_layer(3.0))
    wall = SceneNode(PhongMaterial((0.1,) * 3, (1,) * 3, (1,) * 3, None, None, 10.0), 0, 0, 1.0, 1.0,
                     Isometry3.new((0, 0, 8), (0, 0, 0)), Cuboid((3, 3, 0.1)))
    o, d = np.array([0.0, -0.5, -5.0]), np.array([0.0, 0.0, 1.0])
    for nodes, maxtoi, expect in (([n1], 20.0, "filter1"), ([n1, n2], 20.0, "filter2"), ([n1, n2, wall], 20.0, None),
                                  ([n1, n2, wall], 12.0, "filter2")):
        sc = Scene(nodes, [], upload=False)
        a = O.OracleScene(sc.flat, 64).intersects_ray(o, d, maxtoi)
        b, _ = BF.BruteScene(nodes, []).intersects_ray(o, d, maxtoi)
        if expect is None:
            assert a is None and b is None
            continue
        k = 1 if expect == "filter1" else 2
        want = np.array([0.5, 0.25, 0.125], np.float32) ** k  # ambient colour x (1 - alpha 0), once per node
        assert np.allclose(b, want, rtol=1e-6), (expect, b)
        assert np.allclose(a, want, rtol=1e-6), (expect, a)


def test_exact_ties_keep_one_of_the_tied_nodes():
    """Two coplanar quads in different nodes (exact tie in toi): best_first_search keeps the first found (strict <,
    SURVEY B.2), so the winner is order-dependent — the oracle must return ONE of the tied nodes' colours, never a
    mixture and never the background, and it must do so deterministically."""
    quad = np.array([[-1, -1, 0], [1, -1, 0], [1, 1, 0], [-1, 1, 0]], np.float32)
    Fc = np.array([[0, 1, 2], [0, 2, 3]], np.uint32)
    red = PhongMaterial((1, 0, 0), (0, 0, 0), (0, 0, 0), None, None, 1.0)
    green = PhongMaterial((0, 1, 0), (0, 0, 0), (0, 0, 0), None, None, 1.0)
    nodes = [SceneNode(red, 0, 0, 1.0, 1.0, Isometry3.identity(), TriMesh(quad, Fc, None)),
             SceneNode(green, 0, 0, 1.0, 1.0, Isometry3.identity(), TriMesh(quad.copy(), Fc, None))]
    sc = Scene(nodes, [], upload=False)
    orc = O.OracleScene(sc.flat, 64)
    cam = make_camera(4, 4, 1, 0.0, (0, 0, -3), np.eye(4))
    seen = set()
    for x, y in ((0.1, 0.2), (-0.4, 0.3), (0.5, -0.5)):
        o, d = np.array([x, y, -3.0]), np.array([0.0, 0.0, 1.0])
        c = tuple(np.round(orc.trace(cam, o, d), 6))
        assert c in ((1.0, 0.0, 0.0), (0.0, 1.0, 0.0)), c
        assert tuple(np.round(orc.trace(cam, o, d), 6)) == c
        seen.add(c)
        bi, bh, gap = BF.BruteScene(nodes, []).closest(o, d)
        assert gap == 0.0 and bh.toi == pytest.approx(3.0)


@pytest.mark.parametrize("seed", range(8))
def test_trace_matches_brute_force(seed):
    """Scene::trace for whole (small) images: closest hit, Phong + shadow filter, reflection / refraction recursion and the
    combine of src/scene.rs:171-190.  Pixels whose closest-hit gap or shadow margin is a near-tie are skipped (the
    reference's own answer there depends on its BVT's visit order)."""
    nodes, lights, rng = _rand_scene(3000 + seed)
    sc = Scene(nodes, lights, upload=False)
    w, h = 20, 14
    eye = rng.uniform(-1, 1, 3) + np.array([0, 2.0, -9.0])
    proj = camera_projection(eye, (0, 0, 0), 50.0, w, h)
    cam = make_camera(w, h, 1, 0.0, eye, proj, seed=0, max_depth=6)
    orc = O.OracleScene(sc.flat, 64)
    compared = 0
    for y in range(h):
        for x in range(w):
            bs = BF.BruteScene(nodes, lights, max_depth=6)
            o, d = BF.primary_ray(w, h, eye, proj, x, y)
            ro = O.primary_ray(cam, y * w + x, 0)
            assert np.allclose(ro[0], o) and np.allclose(ro[1], d, atol=1e-12)
            want = bs.trace(o, d)
            if bs.min_gap < 1e-7:
                continue
            got = orc.trace(cam, o, d, y * w + x, 0)
            compared += 1
            assert np.allclose(got, want, rtol=3e-5, atol=3e-5), "pixel (%d,%d): oracle %r brute %r" % (x, y, got, want)
    assert compared > 0.9 * w * h


def fixed_scene():
    """A scene with every ray class (textured reflective floor, refractive box, reflective ball, cone, two lights)."""
    rng = np.random.default_rng(77)
    P, Fc = np.array([[-3, -1, -3], [3, -1, -3], [3, -1, 3], [-3, -1, 3]], np.float32), np.array([[0, 1, 2], [0, 2, 3]], np.uint32)
    UV = np.array([[0, 0], [2, 0], [2, 2], [0, 2]], np.float32)
    tex = _rand_texture(rng, 8, 8, Interpolation.Bilinear, Overflow.Wrap)
    floor = SceneNode(PhongMaterial((0.2,) * 3, (0.8,) * 3, (0.5,) * 3, tex, None, 30.0), 0.3, 0.35, 1.0, 1.0,
                      Isometry3.identity(), TriMesh(P, Fc, UV))
    glass = SceneNode(PhongMaterial((0.1, 0.2, 0.3), (0.5,) * 3, (1,) * 3, None, None, 60.0), 0.0, 0.0, 0.3, 1.4,
                      Isometry3.new((0.8, 0.0, 0.0), (0.2, 0.4, 0.0)), Cuboid((0.6, 0.9, 0.5)))
    ball = SceneNode(PhongMaterial((0.1,) * 3, (1, 0.5, 0.2), (1,) * 3, None, None, 100.0), 0.4, 0.5, 1.0, 1.0,
                     Isometry3.new((-1.0, 0.0, 0.5), (0, 0, 0)), Ball(0.9))
    cone = SceneNode(NormalMaterial(), 0, 0, 1.0, 1.0, Isometry3.new((0.0, -0.2, 1.8), (0.0, 0.0, 0.3)), Cone(0.8, 0.6))
    nodes, lights = [floor, glass, ball, cone], [Light((2, 5, -3), 0.0, 1, (1, 1, 1)), Light((-4, 3, -2), 0.0, 4, (0.4, 0.4, 0.6))]
    return nodes, lights, (0.3, 0.5, 0.9), (0.0, 1.5, -6.0)


def brute_image(nodes, lights, background, eye, w, h, max_depth=8, fovy=45.0):
    proj = camera_projection(eye, (0, 0, 0), fovy, w, h)
    bs = BF.BruteScene(nodes, lights, background, max_depth=max_depth)
    want = np.zeros((w * h, 3), np.float32)
    for y in range(h):
        for x in range(w):
            want[y * w + x] = bs.trace(*BF.primary_ray(w, h, eye, proj, x, y))
    return want, bs, proj


def test_render_counts_and_image_match_brute_force():
    """nro_render (scene::render restated) against the brute force on a fixed scene with every ray class: image and the
    per-class ray counts (reflect / refract / shadow) — also checks the flatten step, which only the oracle goes through."""
    nodes, lights, bg, eye = fixed_scene()
    sc = Scene(nodes, lights, bg, upload=False)
    w, h = 24, 16
    want, bs, proj = brute_image(nodes, lights, bg, eye, w, h)
    cam = make_camera(w, h, 1, 0.0, eye, proj, seed=0, max_depth=8)
    img, st = O.OracleScene(sc.flat, 64).render(cam, threads=2)
    assert bs.min_gap > 1e-7
    assert np.abs(img - want).max() < 5e-5
    assert st.rays_primary == w * h
    assert (st.rays_reflect, st.rays_refract, st.rays_shadow, st.paths_truncated) == (
        bs.counts["reflect"], bs.counts["refract"], bs.counts["shadow"], bs.counts["truncated"])
    assert st.rays_reflect > 0 and st.rays_refract > 0 and st.rays_shadow > 0


@pytest.mark.gpu
def test_device_matches_brute_force_directly(gpu):
    """The CUDA path against the brute-force checker with NO oracle in between: the fixed every-ray-class scene and three
    random scenes (RNG-free: window 0, point lights), through nrb_render."""
    from nrays_b200 import render
    from util import assert_parity

    cases = [fixed_scene()]
    for seed in (3001, 3004, 3006):
        nodes, lights, rng = _rand_scene(seed)
        cases.append((nodes, lights, (1.0, 1.0, 1.0), tuple(rng.uniform(-1, 1, 3) + np.array([0, 2.0, -9.0]))))
    for k, (nodes, lights, bg, eye) in enumerate(cases):
        w, h = 48, 32
        want, bs, proj = brute_image(nodes, lights, bg, eye, w, h, max_depth=6, fovy=50.0)
        scene = Scene(nodes, lights, bg)
        img, st = render(scene, (w, h), 1, 0.0, eye, proj, seed=0, max_depth=6, return_stats=True)
        scene.close()
        assert_parity(img.pixels, want, what="device vs brute force #%d" % k, wh=(w, h))
        for key, name in (("rays_reflect", "reflect"), ("rays_refract", "refract"), ("rays_shadow", "shadow")):
            a, b = int(getattr(st, key)), bs.counts[name]
            assert abs(a - b) <= max(4, 5e-3 * b), (k, key, a, b)
