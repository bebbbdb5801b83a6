"""Semantics of trace / shade / shadow in the oracle, pinned by values derived by hand from the
reference source (file:line in each test)."""
import math

import numpy as np
import pytest

import oracle_lib as O
from nrays_b200 import (Ball, Cuboid, ImageData, Interpolation, Light, NormalMaterial, Overflow, PhongMaterial, Plane,
                        Scene, Texture2d, TriMesh, UVMaterial, make_camera)
from util import checker_texture, default_phong, node, quad_mesh


def flat(nodes, lights, bg=(1.0, 1.0, 1.0)):
    return Scene(nodes, lights, bg, upload=False).flat


def cam(max_depth=0, seed=0):
    return make_camera(4, 4, 1, 0.0, (0, 0, 0), np.eye(4), seed=seed, max_depth=max_depth)


def test_miss_returns_background():
    s = O.OracleScene(flat([node(Ball(1.0), pos=(0, 0, 10))], [], bg=(0.2, 0.4, 0.6)))
    np.testing.assert_allclose(s.trace(cam(), (0, 0, 0), (0, 1, 0)), (0.2, 0.4, 0.6))   # src/scene.rs:169


def test_normal_and_uv_materials():
    s = O.OracleScene(flat([node(Ball(1.0), NormalMaterial(), pos=(0, 0, 5))], []))
    np.testing.assert_allclose(s.trace(cam(), (0, 0, 0), (0, 0, 1)), (0.5, 0.5, 0.0), atol=1e-7)  # (1+n)/2, n=(0,0,-1)
    s = O.OracleScene(flat([node(Ball(1.0), UVMaterial(), pos=(0, 0, 5))], []))
    np.testing.assert_allclose(s.trace(cam(), (0, 0, 0), (0, 0, 1)), (0.25, 0.5, 0.0), atol=1e-7)
    # UVMaterial on a shape without uvs -> (0,0,0,0): alpha 0 -> pure refraction (src/uv_material.rs:18)
    s = O.OracleScene(flat([node(Plane((0, 0, -1)), UVMaterial(), pos=(0, 0, 5))], [], bg=(0.3, 0.3, 0.3)))
    np.testing.assert_allclose(s.trace(cam(), (0, 0, 0), (0, 0, 1)), (0.3, 0.3, 0.3), atol=1e-7)


def test_phong_known_answer_head_on():
    """Light at the eye, head-on hit: ndl = 1, reflected light dir = -ray dir -> scoeff = 1:
    colour = Ka + Lc * (Kd * 1 + Ks * 1^Ns)  (src/phong_material.rs:102-147)."""
    m = PhongMaterial((0.1, 0.2, 0.3), (0.5, 0.4, 0.3), (0.2, 0.1, 0.05), None, None, 30.0)
    s = O.OracleScene(flat([node(Plane((0, 0, -1)), m, pos=(0, 0, 5))], [Light((0, 0, 0), 0.0, 1, (1.0, 0.5, 0.25))]))
    exp = np.array([0.1, 0.2, 0.3]) + np.array([1.0, 0.5, 0.25]) * (np.array([0.5, 0.4, 0.3]) + np.array([0.2, 0.1, 0.05]))
    np.testing.assert_allclose(s.trace(cam(), (0, 0, 0), (0, 0, 1)), exp, atol=1e-6)


def test_phong_specular_not_gated_on_ndl_and_diffuse_clamped():
    """Light behind the surface: diffuse clamps to 0, specular still uses the unclamped ndl."""
    m = PhongMaterial((0, 0, 0), (1, 1, 1), (1, 1, 1), None, None, 2.0)
    L = (0.0, 0.0, 8.0)  # behind the plane z = 5 seen from the origin; the plane itself blocks it
    s = O.OracleScene(flat([node(Plane((0, 0, -1)), m, pos=(0, 0, 5))], [Light(L, 0.0, 1, (1, 1, 1))]))
    # shadow ray starts 0.001 beyond the plane toward the light -> unoccluded; ndl = -1; rl = -l + 2 ndl n
    # = (0,0,-1) + 2*(-1)*(0,0,-1) = (0,0,1); scoeff = -(rl . d) = -1 -> no specular; diffuse = 0
    np.testing.assert_allclose(s.trace(cam(), (0, 0, 0), (0, 0, 1)), (0, 0, 0), atol=1e-7)


def test_shadow_occlusion_and_distance_limit():
    m = default_phong()
    nodes = [node(Plane((0, 1, 0)), m, pos=(0, 0, 0)), node(Ball(0.5), m, pos=(0, 2, 0))]
    s = O.OracleScene(flat(nodes, [Light((0, 4, 0), 0.0, 1, (1, 1, 1))]))
    lit = s.trace(cam(), (3, 1, 0), (0, -1, 0))
    dark = s.trace(cam(), (0.1, 1, 0), (0, -1, 0))
    np.testing.assert_allclose(dark, (0.1, 0.1, 0.1), atol=1e-6)   # ambient only: blocked by the ball
    assert lit[0] > 0.5
    # an occluder beyond the light does not shadow (t.toi <= maxtoi, src/scene.rs:313)
    nodes = [node(Plane((0, 1, 0)), m, pos=(0, 0, 0)), node(Ball(0.5), m, pos=(0, 6, 0))]
    s = O.OracleScene(flat(nodes, [Light((0, 4, 0), 0.0, 1, (1, 1, 1))]))
    assert s.trace(cam(), (0.1, 1, 0), (0, -1, 0))[0] > 0.5


def test_transparent_shadow_filter_per_node():
    """intersects_ray (src/scene.rs:147-161, 304-339): transparent nodes multiply the filter by
    ambient * (1 - alpha); the decision is per SceneNode closest hit."""
    ka = (0.5, 0.25, 1.0)
    glass = PhongMaterial(ka, (1, 1, 1), (0, 0, 0), None, None, 10.0)
    s = O.OracleScene(flat([node(Ball(1.0), glass, pos=(0, 0, 5), alpha=0.2)], []))
    f = s.intersects_ray((0, 0, 0), (0, 0, 1), 100.0)
    np.testing.assert_allclose(f, np.array(ka) * 0.8, atol=1e-7)          # one closest hit of that node only
    assert s.intersects_ray((0, 0, 0), (0, 0, 1), 3.0) is not None         # ball beyond maxtoi: filter (1,1,1)
    np.testing.assert_allclose(s.intersects_ray((0, 0, 0), (0, 0, 1), 3.0), (1, 1, 1))
    # two transparent nodes: product; an opaque node anywhere within maxtoi -> None
    two = [node(Ball(1.0), glass, pos=(0, 0, 5), alpha=0.2), node(Ball(1.0), glass, pos=(0, 0, 9), alpha=0.5)]
    s = O.OracleScene(flat(two, []))
    np.testing.assert_allclose(s.intersects_ray((0, 0, 0), (0, 0, 1), 100.0), np.array(ka) * 0.8 * np.array(ka) * 0.5, atol=1e-7)
    s = O.OracleScene(flat(two + [node(Ball(1.0), default_phong(), pos=(0, 0, 20))], []))
    assert s.intersects_ray((0, 0, 0), (0, 0, 1), 100.0) is None


def test_alpha_mapped_mesh_opaque_triangle_behind_transparent_one_does_not_occlude():
    """F10 / SURVEY A.6: within ONE node only the closest hit counts."""
    w = h = 2
    px = np.ones((4, 4), np.float32)
    px[:, 3] = 0.0  # opacity map: alpha 0 everywhere
    clear = Texture2d(ImageData(px, (w, h)), Interpolation.Nearest, Overflow.ClampToEdges)
    px2 = np.ones((4, 4), np.float32)  # alpha 1 everywhere
    solidmap = Texture2d(ImageData(px2, (w, h)), Interpolation.Nearest, Overflow.ClampToEdges)
    P, F, UV = quad_mesh(2.0, 1, y=0.0)
    P2 = P.copy()
    P2[:, 1] = -1.0
    both = TriMesh(np.concatenate([P, P2]), np.concatenate([F, F + len(P)]), np.concatenate([UV, UV]))
    m = PhongMaterial((1, 1, 1), (1, 1, 1), (0, 0, 0), None, clear, 10.0)
    s = O.OracleScene(flat([node(both, m)], []))
    f = s.intersects_ray((0.3, 3, 0.2), (0, -1, 0), 100.0)
    np.testing.assert_allclose(f, (1, 1, 1))  # closest layer transparent (alpha 0): filter *= Ka * 1; second layer ignored
    m2 = PhongMaterial((1, 1, 1), (1, 1, 1), (0, 0, 0), None, solidmap, 10.0)
    s = O.OracleScene(flat([node(both, m2)], []))
    assert s.intersects_ray((0.3, 3, 0.2), (0, -1, 0), 100.0) is None


def test_reflection_generations_follow_f32_energy():
    """trace_reflection (src/scene.rs:196-218): energy -= attenuation in f32, recurse while energy > 0.1.
    att 0.2: 1, .8, .6000000238, .4000000358, .2000000328, 2.98e-8 -> 5 reflections; att 0.5 -> 2."""
    def count(att):
        mir = [node(Plane((0, 1, 0)), NormalMaterial(), pos=(0, -1, 0), refl=(0.5, att)),
               node(Plane((0, -1, 0)), NormalMaterial(), pos=(0, 1, 0), refl=(0.5, att))]
        sc = O.OracleScene(flat(mir, []))
        # 1x1 image, eye at the origin between the two mirrors; "projection" maps the only pixel's ndc
        # (-1, 1, -1, 1) to the point (-1, -2, -1): a ray that keeps bouncing between y = -1 and y = +1
        P = np.eye(4)
        P[:3, 3] = (0.0, -3.0, 0.0)
        _, st = sc.render(make_camera(1, 1, 1, 0.0, (0, 0, 0), P), 1)
        return st.rays_reflect

    assert count(0.2) == 5
    assert count(0.5) == 2


def test_refraction_direction_and_index_toggle():
    """trace_refraction (src/scene.rs:221-252): nd = normalize(n (d.n) + (d - n (d.n)) * n2/n1), refr toggles."""
    # a transparent slab (cuboid) of NormalMaterial, alpha 0 -> colour = what is behind, through two interfaces
    behind = node(Plane((0, 0, -1)), UVMaterial(), pos=(0, 0, 50))  # never reached colour check; use background
    slab = node(Cuboid((10, 10, 1)), NormalMaterial(), pos=(0, 0, 5), alpha=0.0, refr=1.5)
    s = O.OracleScene(flat([slab], [], bg=(0.25, 0.5, 0.75)))
    d = np.array([0.3, 0.0, 1.0])
    d /= np.linalg.norm(d)
    np.testing.assert_allclose(s.trace(cam(), (0, 0, 0), d), (0.25, 0.5, 0.75), atol=1e-7)  # alpha 0: all weight behind
    half = node(Cuboid((10, 10, 1)), NormalMaterial(), pos=(0, 0, 5), alpha=0.5, refr=1.5)
    s = O.OracleScene(flat([half], [], bg=(0.0, 0.0, 0.0)))
    # entry face normal (0,0,-1): colour ((1+n)/2) = (.5,.5,0), weight alpha = .5; exit (inside hit, inward normal
    # (0,0,-1)) again (.5,.5,0) with weight (1-.5)*.5; then background 0
    np.testing.assert_allclose(s.trace(cam(), (0, 0, 0), d), np.array([0.5, 0.5, 0.0]) * (0.5 + 0.25), atol=1e-6)


def test_depth_cap_counts_truncations():
    mir = [node(Plane((0, 1, 0)), NormalMaterial(), pos=(0, -1, 0), refl=(0.5, 0.0)),
           node(Plane((0, -1, 0)), NormalMaterial(), pos=(0, 1, 0), refl=(0.5, 0.0))]
    sc = O.OracleScene(flat(mir, []))
    P = np.eye(4)
    P[:3, 3] = (0.0, -3.0, 0.0)
    c = make_camera(1, 1, 1, 0.0, (0, 0, 0), P, max_depth=7)
    _, st = sc.render(c, 1)
    assert st.rays_reflect == 6 and st.paths_truncated == 1   # depths 0..6 traced, the 7th spawn is cut


def numpy_sample(tex, u, v):
    """Independent numpy restatement of Texture2d::sample (src/texture2d.rs:207-256)."""
    w, h = tex.data.dims
    px = tex.data.pixels
    ux, uy = np.float32(u), np.float32(v)
    if tex.overflow == Overflow.ClampToEdges:
        ux, uy = np.clip(ux, np.float32(0), np.float32(1)), np.clip(uy, np.float32(0), np.float32(1))
    else:
        ux, uy = np.float32(math.fmod(ux, 1.0)), np.float32(math.fmod(uy, 1.0))
        if ux < 0:
            ux = np.float32(1) + ux
        if uy < 0:
            uy = np.float32(1) + uy
    ux, uy = ux * np.float32(w - 1), uy * np.float32(h - 1)

    def at(x, y):
        return px[min(int(y) * w + int(x), len(px) - 1)]

    if tex.interpol == Interpolation.Nearest:
        return at(np.floor(ux + np.float32(0.5)) if ux >= 0 else ux, np.floor(uy + np.float32(0.5)))
    lx, ly = np.floor(ux), np.floor(uy)
    sx, sy = ux - lx, uy - ly
    up = at(lx, ly + 1) * (np.float32(1) - sx) + at(lx + 1, ly + 1) * sx
    dn = at(lx, ly) * (np.float32(1) - sx) + at(lx + 1, ly) * sx
    return up * sy + dn * (np.float32(1) - sy)


@pytest.mark.parametrize("interp", [Interpolation.Bilinear, Interpolation.Nearest])
@pytest.mark.parametrize("overflow", [Overflow.Wrap, Overflow.ClampToEdges])
def test_texture_sample_matches_numpy(interp, overflow):
    tex = Texture2d(checker_texture(7, 5), interp, overflow)
    m = PhongMaterial((1, 1, 1), (1, 1, 1), (0, 0, 0), tex, None, 1.0)
    s = O.OracleScene(flat([node(Ball(1.0), m)], []))
    rng = np.random.default_rng(2)
    for u, v in list(rng.uniform(-2.5, 2.5, (200, 2))) + [(0.0, 0.0), (1.0, 1.0), (0.999999, 0.5), (-0.25, 1.75)]:
        np.testing.assert_allclose(s.texture_sample(0, u, v), numpy_sample(tex, u, v), rtol=0, atol=1e-6)


def test_from_array_decode_conventions():
    """Texture2d::from_png channel expansion (src/texture2d.rs:96-173): y flip; depth 3 opacity = red; depth 4 diffuse drops alpha."""
    img = np.zeros((2, 1, 4), np.uint8)
    img[0, 0] = (255, 0, 0, 51)
    img[1, 0] = (0, 255, 0, 102)
    d = Texture2d.from_array(img, False, 0, 0).data.pixels
    np.testing.assert_allclose(d, [[0, 1, 0, 1], [1, 0, 0, 1]])         # flipped, alpha forced to 1
    o = Texture2d.from_array(img, True, 0, 0).data.pixels
    np.testing.assert_allclose(o, [[1, 1, 1, 0.4], [1, 1, 1, 0.2]], atol=1e-7)
    o3 = Texture2d.from_array(img[:, :, :3], True, 0, 0).data.pixels
    np.testing.assert_allclose(o3, [[1, 1, 1, 0], [1, 1, 1, 1]])        # depth 3 opacity reads RED
    g = Texture2d.from_array(img[:, :, 1], False, 0, 0).data.pixels
    np.testing.assert_allclose(g, [[1, 1, 1, 1], [0, 0, 0, 1]])


def test_primary_ray_matches_closed_form():
    """A.1/A.2: unprojecting ndc through (P V)^-1 equals eye + f + ndc.x*aspect*t*r + ndc.y*t*u."""
    from nrays_b200 import camera_projection
    eye, at, fovy, w, h = np.array([3.0, 2.0, -7.0]), np.array([0.5, 0.0, 1.0]), 38.0, 64, 48
    c = make_camera(w, h, 1, 0.0, eye, camera_projection(eye, at, fovy, w, h))
    f = (at - eye) / np.linalg.norm(at - eye)
    r = np.cross(f, (0, 1, 0))
    r /= np.linalg.norm(r)
    u = np.cross(r, f)
    t = math.tan(math.radians(fovy) / 2)
    for pix in (0, 17, w * h - 1, w * 10 + 5):
        y, x = divmod(pix, w)
        ndx, ndy = (x / w - 0.5) * 2, -(y / h - 0.5) * 2     # no pixel-centre offset
        exp = f + ndx * (w / h) * t * r + ndy * t * u
        exp /= np.linalg.norm(exp)
        o, d = O.primary_ray(c, pix)
        np.testing.assert_allclose(o, eye)
        np.testing.assert_allclose(d, exp, atol=1e-9)


def test_jitter_is_seeded_and_bounded():
    from nrays_b200 import camera_projection
    P = camera_projection((0, 0, -5), (0, 0, 0), 45, 32, 32)
    a = O.primary_ray(make_camera(32, 32, 4, 1.0, (0, 0, -5), P, seed=1), 100, 2)[1]
    b = O.primary_ray(make_camera(32, 32, 4, 1.0, (0, 0, -5), P, seed=1), 100, 2)[1]
    c = O.primary_ray(make_camera(32, 32, 4, 1.0, (0, 0, -5), P, seed=2), 100, 2)[1]
    z = O.primary_ray(make_camera(32, 32, 4, 0.0, (0, 0, -5), P, seed=1), 100, 2)[1]
    np.testing.assert_array_equal(a, b)
    assert np.abs(a - c).max() > 0
    assert 0 < np.abs(a - z).max() < 0.05   # within half a pixel of the unjittered ray
