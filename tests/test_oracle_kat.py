"""Known-answer tests that pin the oracle (the reference ships none: SURVEY.md §4, §8c).

Closed-form values are computed by hand / by an independent numeric method, never by the oracle."""
import math

import numpy as np
import pytest

import oracle_lib as O
from nrays_b200 import (Ball, Capsule, Cone, Cuboid, Cylinder, Isometry3, Light, NormalMaterial, Plane, Scene, SceneNode,
                        TriMesh)


def one(geom, pos=(0, 0, 0), angle=(0, 0, 0), solid=False, bits=64):
    n = SceneNode(NormalMaterial(), 0, 0, 1.0, 1.0, Isometry3.new(pos, np.radians(angle)), geom, None, solid)
    return O.OracleScene(Scene([n], [], upload=False).flat, bits)


# ---- Philox4x32-10: Random123 known-answer vectors (kat_vectors) ---------------------------------
@pytest.mark.parametrize("ctr,key,exp", [
    ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
    ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
    ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
     [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
])
def test_philox_known_answers(ctr, key, exp):
    assert O.philox(ctr, key) == exp


# ---- AABB::toi_with_ray --------------------------------------------------------------
def test_aabb_slab_cases():
    lo, hi = (-1, -1, -1), (1, 1, 1)
    assert O.aabb_toi(lo, hi, (0, 0, -5), (0, 0, 1)) == pytest.approx(4.0)
    assert O.aabb_toi(lo, hi, (0, 0, 0), (0, 0, 1), solid=True) == 0.0          # origin inside, solid
    assert O.aabb_toi(lo, hi, (0, 0, 0), (0, 0, 1), solid=False) == pytest.approx(1.0)  # exit distance
    assert O.aabb_toi(lo, hi, (2, 0, -5), (0, 0, 1)) is None                    # zero dir component outside the slab
    assert O.aabb_toi(lo, hi, (1, 0, -5), (0, 0, 1)) == pytest.approx(4.0)      # zero dir component on the face: inside
    assert O.aabb_toi(lo, hi, (0, 0, 5), (0, 0, 1)) is None                     # behind the ray
    d = np.array([1.0, 1.0, 1.0]) / math.sqrt(3)
    assert O.aabb_toi(lo, hi, (-3, -3, -3), d) == pytest.approx(2 * math.sqrt(3))


# ---- Ball ------------------------------------------------------------------------------
def test_ball_known_answer():
    s = one(Ball(1.0))
    r = s.cast(0, (0, 0, -5), (0, 0, 1))
    assert r["toi"] == pytest.approx(4.0)
    np.testing.assert_allclose(r["normal"], (0, 0, -1), atol=1e-12)
    np.testing.assert_allclose(r["uv"], (0.25, 0.5), atol=1e-12)  # u = .5 + atan2(-1, 0)/2pi, v = .5 - asin(0)/pi
    assert s.cast(0, (0, 2, -5), (0, 0, 1)) is None
    assert s.cast(0, (0, 0, 5), (0, 0, 1)) is None  # c > 0 and b > 0


def test_ball_inside_solid_and_hollow():
    r = one(Ball(2.0)).cast(0, (0, 0, 0), (1, 0, 0))
    assert r["toi"] == pytest.approx(2.0)
    np.testing.assert_allclose(r["normal"], (-1, 0, 0), atol=1e-12)  # negated when inside
    np.testing.assert_allclose(r["uv"], (0.5, 0.5), atol=1e-12)      # uv from the un-negated normal
    r = one(Ball(2.0), solid=True).cast(0, (0.5, 0, 0), (1, 0, 0))
    assert r["toi"] == 0.0


def test_ball_ignores_rotation_uses_translation():
    a = one(Ball(1.0), pos=(1, 2, 3), angle=(30, 40, 50)).cast(0, (1, 2, -5), (0, 0, 1))
    assert a["toi"] == pytest.approx(7.0)
    np.testing.assert_allclose(a["uv"], (0.25, 0.5), atol=1e-12)


# ---- Cuboid ----------------------------------------------------------------------------
def test_cuboid_faces_normals_uvs():
    s = one(Cuboid((1, 2, 3)))
    r = s.cast(0, (-5, 0.5, 1.0), (1, 0, 0))
    assert r["toi"] == pytest.approx(4.0)
    np.testing.assert_allclose(r["normal"], (-1, 0, 0))
    np.testing.assert_allclose(r["uv"], ((0.5 + 2) / 4, (1.0 + 3) / 6))  # x face: (dpt.y/sy, dpt.z/sz)
    r = s.cast(0, (0.25, 7, 0.0), (0, -1, 0))
    np.testing.assert_allclose(r["normal"], (0, 1, 0))
    np.testing.assert_allclose(r["uv"], (3 / 6, 1.25 / 2))               # y face: (dpt.z/sz, dpt.x/sx)
    r = s.cast(0, (0.5, -1, 9), (0, 0, -1))
    np.testing.assert_allclose(r["normal"], (0, 0, 1))
    np.testing.assert_allclose(r["uv"], (1.5 / 2, 1 / 4))                # z face: (dpt.x/sx, dpt.y/sy)


def test_cuboid_inside_exit_normal_points_against_ray():
    r = one(Cuboid((1, 1, 1))).cast(0, (0, 0, 0), (0, 1, 0))
    assert r["toi"] == pytest.approx(1.0)
    np.testing.assert_allclose(r["normal"], (0, -1, 0))
    r = one(Cuboid((1, 1, 1)), solid=True).cast(0, (0, 0, 0), (0, 1, 0))
    assert r["toi"] == 0.0


def test_cuboid_rotated_transform():
    s = one(Cuboid((1, 1, 1)), pos=(0, 0, 0), angle=(0, 45, 0))
    r = s.cast(0, (0, 0, -5), (0, 0, 1))
    assert r["toi"] == pytest.approx(5 - math.sqrt(2))
    assert abs(np.linalg.norm(r["normal"]) - 1) < 1e-12 and r["normal"][2] < 0


# ---- Plane -----------------------------------------------------------------------------
def test_plane_both_sides_and_parallel():
    s = one(Plane((0, 1, 0)), pos=(0, -3, 0))
    r = s.cast(0, (0, 5, 0), (0, -1, 0))
    assert r["toi"] == pytest.approx(8.0) and r["uv"] is None
    np.testing.assert_allclose(r["normal"], (0, 1, 0))
    r = s.cast(0, (0, -5, 0), (0, 1, 0))        # from behind: normal flipped toward the origin
    assert r["toi"] == pytest.approx(2.0)
    np.testing.assert_allclose(r["normal"], (0, -1, 0))
    assert s.cast(0, (0, 5, 0), (0, 1, 0)) is None      # pointing away
    assert s.cast(0, (0, 5, 0), (1, 0, 0)) is None      # parallel (t = -inf)
    assert s.cast(0, (0, -5, 0), (1, 0, 0)) is None     # parallel from behind: t = +inf fails cost < f64::MAX
    assert one(Plane((0, 1, 0)), solid=True).cast(0, (0, -1, 0), (0, -1, 0))["toi"] == 0.0


# ---- Triangle / TriMesh ------------------------------------------------------------------
def tri_scene(order=(0, 1, 2)):
    P = np.array([[0, 0, 0], [2, 0, 0], [0, 2, 0]], np.float32)
    UV = np.array([[0, 0], [1, 0], [0, 1]], np.float32)
    return one(TriMesh(P, np.array([order], np.uint32), UV))


@pytest.mark.parametrize("order", [(0, 1, 2), (0, 2, 1)])
def test_triangle_two_sided_normal_faces_origin(order):
    s = tri_scene(order)
    r = s.cast(0, (0.5, 0.5, 3), (0, 0, -1))
    assert r["toi"] == pytest.approx(3.0)
    np.testing.assert_allclose(r["normal"], (0, 0, 1), atol=1e-12)
    np.testing.assert_allclose(r["uv"], (0.25, 0.25), atol=1e-12)  # barycentric blend of vertex uvs
    r = s.cast(0, (0.5, 0.5, -3), (0, 0, 1))
    assert r["toi"] == pytest.approx(3.0)
    np.testing.assert_allclose(r["normal"], (0, 0, -1), atol=1e-12)
    assert s.cast(0, (1.5, 1.5, 3), (0, 0, -1)) is None            # outside the hypotenuse
    assert s.cast(0, (0.5, 0.5, 3), (0, 0, 1)) is None             # pointing away
    assert s.cast(0, (0.5, 0.5, 3), (1, 0, 0)) is None             # parallel to the plane


def test_trimesh_applies_isometry_and_picks_closest():
    P, F = np.array([[-1, -1, 0], [1, -1, 0], [0, 1, 0], [-1, -1, 1], [1, -1, 1], [0, 1, 1]], np.float32), \
        np.array([[0, 1, 2], [3, 4, 5]], np.uint32)
    s = one(TriMesh(P, F, None), pos=(0, 0, 10), angle=(0, 0, 90))
    r = s.cast(0, (0, 0, 0), (0, 0, 1))
    assert r["toi"] == pytest.approx(10.0)
    r = s.cast(0, (0, 0, 20), (0, 0, -1))
    assert r["toi"] == pytest.approx(9.0)
    np.testing.assert_allclose(r["uv"], (0, 0), atol=1e-12)  # no uvs -> zero filled (src/obj.rs:383)


# ---- Cylinder / Cone / Capsule: closed forms checked against an independent dense march -------------
def inside_cylinder(p, hh, r):
    return abs(p[1]) <= hh and p[0] ** 2 + p[2] ** 2 <= r * r


def inside_cone(p, hh, r):
    s = hh - p[1]
    return 0 <= s <= 2 * hh and p[0] ** 2 + p[2] ** 2 <= (r * s / (2 * hh)) ** 2


def inside_capsule(p, hh, r):
    cy = min(max(p[1], -hh), hh)
    return p[0] ** 2 + (p[1] - cy) ** 2 + p[2] ** 2 <= r * r


def march(inside, o, d, want_inside, tmax=30.0, n=6000):
    """First t where the point's inside-state equals want_inside (dense sampling + bisection)."""
    ts = np.linspace(0, tmax, n)
    prev = 0.0
    for t in ts[1:]:
        if inside(o + d * t) == want_inside:
            a, b = prev, t
            for _ in range(60):
                m = 0.5 * (a + b)
                if inside(o + d * m) == want_inside:
                    b = m
                else:
                    a = m
            return b
        prev = t
    return None


@pytest.mark.parametrize("geom,inside", [
    (Cylinder(1.0, 0.7), lambda p: inside_cylinder(p, 1.0, 0.7)),
    (Cone(1.0, 1.0), lambda p: inside_cone(p, 1.0, 1.0)),
    (Capsule(0.8, 0.5), lambda p: inside_capsule(p, 0.8, 0.5)),
])
def test_revolution_solids_match_dense_march(geom, inside):
    s = one(geom)
    rng = np.random.default_rng(5)
    hits = misses = 0
    for _ in range(300):
        o = rng.normal(size=3)
        o = o / np.linalg.norm(o) * rng.uniform(2.5, 5.0)
        target = rng.uniform(-1.2, 1.2, 3)
        d = target - o
        d /= np.linalg.norm(d)
        exp = march(inside, o, d, True)
        got = s.cast(0, o, d)
        if exp is None:
            if got is not None:  # the march can miss grazing slivers; a reported hit must then be very thin
                pm = o + d * (got["toi"] + 1e-3)
                assert not inside(o + d * (got["toi"] + 0.02)) or inside(pm)
            misses += 1
            continue
        assert got is not None, (o, d, exp)
        assert got["toi"] == pytest.approx(exp, abs=2e-6)
        assert got["uv"] is None
        p = o + d * got["toi"]
        # outward unit normal: stepping along it leaves the solid, against it enters
        assert abs(np.linalg.norm(got["normal"]) - 1) < 1e-9
        assert not inside(p + got["normal"] * 1e-4)
        hits += 1
    assert hits > 100 and misses > 10


@pytest.mark.parametrize("geom,inside", [
    (Cylinder(1.0, 0.7), lambda p: inside_cylinder(p, 1.0, 0.7)),
    (Cone(1.0, 1.0), lambda p: inside_cone(p, 1.0, 1.0)),
    (Capsule(0.8, 0.5), lambda p: inside_capsule(p, 0.8, 0.5)),
])
def test_revolution_solids_origin_inside(geom, inside):
    """Origin inside, !solid -> exit point with the OUTWARD normal; solid -> toi 0."""
    rng = np.random.default_rng(9)
    hollow, solid = one(geom), one(geom, solid=True)
    n = 0
    while n < 100:
        o = rng.uniform(-0.6, 0.6, 3)
        if not inside(o):
            continue
        d = rng.normal(size=3)
        d /= np.linalg.norm(d)
        exp = march(inside, o, d, False, tmax=6.0)
        got = hollow.cast(0, o, d)
        assert got["toi"] == pytest.approx(exp, abs=2e-6)
        assert np.dot(got["normal"], d) > 0  # outward = along the exiting ray
        assert solid.cast(0, o, d)["toi"] == 0.0
        n += 1


def test_f32_twin_agrees_with_f64_on_primitives():
    rng = np.random.default_rng(3)
    for geom in (Ball(1.0), Cuboid((1, 0.5, 0.8)), Cylinder(1, 0.6), Cone(1, 1), Capsule(0.7, 0.4)):
        a, b = one(geom, pos=(0.3, -0.2, 0.1), angle=(10, 20, 30)), one(geom, pos=(0.3, -0.2, 0.1), angle=(10, 20, 30), bits=32)
        for _ in range(100):
            o = rng.normal(size=3)
            o = o / np.linalg.norm(o) * 4
            d = rng.uniform(-0.8, 0.8, 3) - o
            d /= np.linalg.norm(d)
            ra, rb = a.cast(0, o, d), b.cast(0, o, d)
            if ra is None or rb is None:
                continue
            assert ra["toi"] == pytest.approx(rb["toi"], abs=2e-4)


def test_light_racsample_is_floor_sqrt():
    # ((nsample as f32).sqrt()) as usize — src/light.rs:20 (nsample 10 -> 3, 50 -> 7)
    assert [Light((0, 0, 0), 0, n, (1, 1, 1)).racsample for n in (1, 3, 4, 10, 50, 100)] == [1, 1, 2, 3, 7, 10]


# ---- SceneNode.nmap: the depth shift of src/scene_node.rs:60-70 -------------------------------------
def _const_tex(rgb):
    from nrays_b200 import ImageData, Interpolation, Overflow, Texture2d

    px = np.ones((4, 4), np.float32)
    px[:, :3] = rgb
    return Texture2d(ImageData(px, (2, 2)), Interpolation.Bilinear, Overflow.Wrap)


def test_nmap_shifts_toi_by_the_mean_texel_and_only_where_the_cast_has_uvs():
    from nrays_b200 import make_camera

    def scene(nodes):
        return O.OracleScene(Scene(nodes, [], upload=False).flat, 64)

    def n(geom, nmap, pos=(0, 0, 0), mat=None):
        return SceneNode(mat or NormalMaterial(), 0, 0, 1.0, 1.0, Isometry3.new(pos, (0, 0, 0)), geom, nmap, False)

    # unit ball: toi 4 -> 4 - (0.3 + 0.6 + 0.9) / 3 = 3.4; normal and uv are those of the real hit
    r = scene([n(Ball(1.0), _const_tex((0.3, 0.6, 0.9)))]).cast(0, (0, 0, -5), (0, 0, 1))
    assert r["toi"] == pytest.approx(4.0 - 0.6, abs=1e-6)
    np.testing.assert_allclose(r["normal"], (0, 0, -1), atol=1e-12)
    np.testing.assert_allclose(r["uv"], (0.25, 0.5), atol=1e-12)
    # cylinder casts return no uvs: the texture is inert
    r = scene([n(Cylinder(1.0, 1.0), _const_tex((0.9, 0.9, 0.9)))]).cast(0, (0, 0, -5), (0, 0, 1))
    assert r["toi"] == pytest.approx(4.0)
    # best-first order: the far ball's shifted toi (7 - 3.5 = 3.5... here 6 - 0.9 = 5.1) would beat nothing; a far
    # node whose AABB starts behind the current best is never cast, even though its shifted toi would win
    cam = make_camera(4, 4, 1, 0.0, (0, 0, -5), np.eye(4), seed=0)
    from nrays_b200 import UVMaterial
    near = n(Ball(1.0), None)
    far_small = n(Ball(1.0), _const_tex((0.5, 0.5, 0.5)), (0, 0, 2.2), UVMaterial())
    far_big = n(Ball(1.0), _const_tex((3.0, 3.0, 3.0)), (0, 0, 2.2), UVMaterial())   # shifted toi 6.2 - 3 = 3.2 < 4
    # near ball: toi 4, normal (0,0,-1) -> NormalMaterial colour (.5,.5,0).  far ball (uv colour (.25,.5,0)): AABB entry
    # 6.2 >= 4 -> pruned by the search, whatever its shift
    for far in (far_small, far_big):
        rgb = scene([near, far]).trace(cam, (0, 0, -5), (0, 0, 1))
        np.testing.assert_allclose(rgb, (0.5, 0.5, 0.0), atol=1e-6)
    # ... but a far node whose AABB starts before the best cost IS cast and wins with its shifted toi
    rgb = scene([near, n(Ball(1.0), _const_tex((0.5, 0.5, 0.5)), (0, 0, -0.2), UVMaterial())]).trace(cam, (0, 0, -5), (0, 0, 1))
    np.testing.assert_allclose(rgb, (0.25, 0.5, 0.0), atol=1e-6)
    # overlapping boxes: far ball at z = 0.5 has AABB entry 4.5 >= 4 -> still pruned; at z = -0.2 (entry 3.8 < 4) it is cast,
    # its real toi is 3.8, shifted 3.8 - 0.5 = 3.3 < 4 -> it wins; its normal is (0,0,-1) too, so compare through toi
    s = scene([near, n(Ball(1.0), _const_tex((0.5, 0.5, 0.5)), (0, 0, -0.2))])
    assert s.cast(1, (0, 0, -5), (0, 0, 1))["toi"] == pytest.approx(3.3, abs=1e-6)
    # shadow query: the shifted toi is what is compared with maxtoi (src/scene.rs:313)
    s = scene([n(Ball(1.0), _const_tex((0.6, 0.6, 0.6)))])
    assert s.intersects_ray((0, 0, -5), (0, 0, 1), 3.5) is None          # 4 - 0.6 = 3.4 <= 3.5: occluded
    assert s.intersects_ray((0, 0, -5), (0, 0, 1), 3.3) is not None      # 3.4 > 3.3: free
