"""Shared helpers for the parity tests."""
import numpy as np

import oracle_lib as O
from nrays_b200 import (Ball, Capsule, Cone, Cuboid, Cylinder, ImageData, Interpolation, Isometry3, Light,
                        NormalMaterial, Overflow, PhongMaterial, Plane, Scene, SceneNode, Texture2d, TriMesh, UVMaterial,
                        camera_projection, make_camera, render)

# The stated float tolerance of the path (BASELINE.md §4 / north_star): per-channel |delta| <= 1/255 on
# >= 99.9 % of the pixels; the remaining pixels are silhouette / shadow-edge flips between the f64
# reference arithmetic and the f32 device arithmetic.
TOL = 1.0 / 255.0
MAX_FRAC_OVER = 1.0e-3


def image_metrics(a, b):
    a = np.asarray(a, dtype=np.float64).reshape(-1, 3)
    b = np.asarray(b, dtype=np.float64).reshape(-1, 3)
    d = np.abs(a - b).max(axis=1)
    return dict(max_abs=float(d.max()) if len(d) else 0.0, mean_abs=float(np.abs(a - b).mean()) if len(d) else 0.0,
                frac_over=float((d > TOL).mean()) if len(d) else 0.0)


def assert_parity(a, b, max_frac=MAX_FRAC_OVER, what=""):
    m = image_metrics(a, b)
    assert np.isfinite(np.asarray(a)).all(), "non-finite pixels " + what
    assert m["frac_over"] <= max_frac, "parity %s: %r" % (what, m)
    return m


def checker_texture(w=16, h=8, seed=1):
    rng = np.random.default_rng(seed)
    px = rng.uniform(0.0, 1.0, (h * w, 4)).astype(np.float32)
    return ImageData(px, (w, h))


def quad_mesh(size=2.0, n=4, y=0.0):
    """An n x n grid of quads in the XZ plane (2 n^2 triangles) with uvs."""
    s = np.linspace(-size, size, n + 1)
    X, Z = np.meshgrid(s, s, indexing="ij")
    P = np.stack([X, np.full_like(X, y), Z], -1).reshape(-1, 3).astype(np.float32)
    UV = np.stack([(X + size) / (2 * size), (Z + size) / (2 * size)], -1).reshape(-1, 2).astype(np.float32)
    idx = np.arange((n + 1) * (n + 1)).reshape(n + 1, n + 1)
    a, b, c, d = idx[:-1, :-1], idx[1:, :-1], idx[1:, 1:], idx[:-1, 1:]
    F = np.concatenate([np.stack([a, b, c], -1).reshape(-1, 3), np.stack([a, c, d], -1).reshape(-1, 3)]).astype(np.uint32)
    return P, F, UV


def default_phong():
    return PhongMaterial((0.1, 0.1, 0.1), (1.0, 1.0, 1.0), (1.0, 1.0, 1.0), None, None, 100.0)


def node(geom, mat=None, pos=(0, 0, 0), angle=(0, 0, 0), refl=(0.0, 0.0), alpha=1.0, refr=1.0, solid=False):
    mat = mat or default_phong()
    return SceneNode(mat, refl[0], refl[1], alpha, refr, Isometry3.new(pos, np.radians(angle)), geom, None, solid)


def look(eye, at, fovy, w, h):
    return camera_projection(eye, at, fovy, w, h)


def render_both(nodes, lights, eye, at=(0, 0, 0), fovy=45.0, w=96, h=64, spp=1, window=0.0, seed=0, bits=64,
                background=(1.0, 1.0, 1.0), max_depth=0):
    """Render the same flattened scene on the device (through the C-ABI) and on the oracle."""
    scene = Scene(nodes, lights, background)
    proj = look(eye, at, fovy, w, h)
    img, st = render(scene, (w, h), spp, window, eye, proj, seed=seed, max_depth=max_depth, return_stats=True)
    cam = make_camera(w, h, spp, window, eye, proj, seed=seed, max_depth=max_depth)
    ref, ost = O.OracleScene(scene.flat, bits).render(cam)
    scene.close()
    return img.pixels, st, ref, ost


def assert_counts_close(st, ost, rel=2e-3):
    for k in ("rays_primary", "rays_reflect", "rays_refract", "rays_shadow", "paths_truncated"):
        a, b = int(getattr(st, k)), int(getattr(ost, k))
        assert abs(a - b) <= max(4, rel * max(a, b)), "%s: device %d vs oracle %d" % (k, a, b)
    assert st.rays_primary == ost.rays_primary
