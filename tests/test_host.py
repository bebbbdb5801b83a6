"""Host-side logic above the C-ABI: the .scene / .mtl parser twin (examples/loader3d.rs, src/mtl.rs),
camera matrices (nalgebra), flattening into the POD tables, tile packing."""
import math

import numpy as np
import pytest

from nrays_b200 import (Ball, Isometry3, Light, NormalMaterial, PhongMaterial, Scene, SceneNode, Texture2d, TriMesh,
                        _abi as A, camera_projection, configs, dist, perspective3)
from nrays_b200.loader3d import AssetResolver, Camera, ObjData, SceneFileError, parse, parse_mtl

SCENE = """
# a comment
mtllib lib.mtl
camera
  output o.png
  resolution 640 480
  eye 0 1 -5
  at 0 0 0
  fovy 60
light
  pos 1 2 3
  color 1 0.5 0.25
  radius 0.5
  nsample 50
geometry
  ball 2.5
  cone 1 1
  pos 1 0 0
  angle 0 90 0
  material shiny
  refl 0.3 0.25
  refr 1.4
  solid
geometry
  plane 0 2 0
  pos 0 -1 0
  angle 0 0 0
  material normals
unknown line is ignored
"""
MTL = """
newmtl shiny
Ka 0.1 0.2 0.3
Kd 0.4 0.5 0.6
Ks 0.7 0.8 0.9
Ns 12
d 0.5
map_Kd tex.png
map_d mask.png
newmtl plain
"""


def resolver():
    r = AssetResolver()
    r.files["lib.mtl"] = MTL
    r.textures["tex.png"] = np.full((2, 3, 3), 128, np.uint8)
    r.textures["mask.png"] = np.full((2, 3), 255, np.uint8)
    return r


def test_parse_scene_file_grammar_and_defaults():
    lights, nodes, cameras = parse(SCENE, resolver())
    assert len(lights) == 1 and len(nodes) == 2 and len(cameras) == 1
    c = cameras[0]
    assert c.resolution == (640.0, 480.0) and c.fovy == 60.0 and c.aa == (1.0, 0.0)     # aa default: loader3d.rs:426
    l = lights[0]
    assert l.racsample == 7 and l.radius == 0.5 and l.color == (1.0, 0.5, 0.25)         # floor(sqrt(50))
    n = nodes[0]
    assert n.geometry.kind == A.NRB_SHAPE_BALL and n.geometry.param[0] == 2.5            # only the first shape (F11)
    assert n.refl_mix == pytest.approx(0.3) and n.refl_atenuation == 0.25 and n.refr_coeff == 1.4 and n.solid
    assert n.alpha == 0.5 and n.material.shininess == 12.0
    assert n.material.texture is not None and n.material.alpha is not None
    np.testing.assert_allclose(n.transform.rot @ np.array([1, 0, 0]), (0, 0, -1), atol=1e-12)  # axis-angle 90 deg about y
    p = nodes[1]
    assert p.geometry.kind == A.NRB_SHAPE_PLANE and p.geometry.param == (0.0, 1.0, 0.0)  # normalised
    assert isinstance(p.material, NormalMaterial) and p.refl_mix == 0.0 and p.refr_coeff == 1.0 and p.alpha == 1.0


def test_mtl_defaults():
    ms = parse_mtl(MTL)
    assert [m.name for m in ms] == ["shiny", "plain"]
    plain = ms[1]
    assert plain.shininess == 60.0 and plain.alpha == 1.0 and plain.ambiant == (1.0, 1.0, 1.0)  # src/mtl.rs:149-162


@pytest.mark.parametrize("text,msg", [
    ("camera\n output o.png\n eye 0 0 0\n at 0 0 1\n fovy 45\n", "resolution"),
    ("light\n pos 0 0 0\n", "color"),
    ("geometry\n ball 1\n pos 0 0 0\n angle 0 0 0\n material nope\n", "unknown material"),
    ("geometry\n pos 0 0 0\n angle 0 0 0\n material default\n", "geom_type"),
    ("geometry\n ball 1\n pos 0 0\n", "3 components"),
    ("camera\n output o\n resolution 4 4\n eye 0 0 0\n at 0 0 1\n fovy 4\n aa 0 1\n", "at least 1"),
])
def test_parse_errors_where_the_reference_panics(text, msg):
    with pytest.raises(SceneFileError) as e:
        parse(text, resolver())
    assert msg in str(e.value)


def test_obj_groups_share_vertices_and_scale_by_quarter():
    P = np.array([[0, 0, 0], [4, 0, 0], [0, 4, 0], [0, 0, 4]], np.float32)
    od = ObjData(P, None, [("a", np.array([[0, 1, 2]], np.uint32), None), ("b", np.array([[0, 2, 3]], np.uint32), None),
                           ("empty", np.zeros((0, 3), np.uint32), None)])
    r = AssetResolver()
    r.objs["m.obj"] = od
    text = "geometry\n obj m.obj dir\n pos 0 0 0\n angle 0 0 0\n material default\n"
    _l, nodes, _c = parse(text, r)
    assert len(nodes) == 2                                               # one SceneNode per non-empty group
    np.testing.assert_allclose(nodes[0].geometry.coords[1], (1, 0, 0))   # /4: loader3d.rs:669
    flat = Scene(nodes, [], upload=False).flat
    assert len(flat.positions) == 4 and flat.n_triangles == 2            # shared vertex array stored once
    assert flat.node_rows[0].vertex_base == flat.node_rows[1].vertex_base == 0
    assert flat.node_rows[1].first_index == 3


def test_perspective_and_look_at_match_nalgebra_formulas():
    P = perspective3(16 / 9, math.radians(45), 1.0, 100000.0)
    t = math.tan(math.radians(45) / 2)
    assert P[0, 0] == pytest.approx(1 / (16 / 9 * t)) and P[1, 1] == pytest.approx(1 / t)
    assert P[2, 2] == pytest.approx((100000.0 + 1) / (1 - 100000.0)) and P[2, 3] == pytest.approx(2 * 100000.0 / (1 - 100000.0))
    assert P[3, 2] == -1.0
    V = Isometry3.look_at_rh((1, 2, 3), (4, 5, 6), (0, 1, 0)).to_homogeneous()
    np.testing.assert_allclose(V @ np.array([1, 2, 3, 1.0]), (0, 0, 0, 1), atol=1e-12)        # eye -> origin
    f = V @ np.array([4, 5, 6, 1.0])
    assert f[2] < 0 and abs(f[0]) < 1e-12 and abs(f[1]) < 1e-12                                # looks down -z
    M = camera_projection((1, 2, 3), (4, 5, 6), 45, 1920, 1080)
    np.testing.assert_allclose(M @ (P @ V), np.eye(4), atol=1e-6)


def test_flatten_dedups_textures_and_materials():
    data = Texture2d.from_array(np.zeros((2, 2, 3), np.uint8), False, 0, 1).data
    t1, t2 = Texture2d(data, 0, 1), Texture2d(data, 0, 1)
    m = PhongMaterial((1, 1, 1), (1, 1, 1), (1, 1, 1), t1, None, 5.0)
    m2 = PhongMaterial((1, 1, 1), (1, 1, 1), (1, 1, 1), t2, None, 5.0)
    nodes = [SceneNode(m, 0, 0, 1, 1, Isometry3.identity(), Ball(1.0)), SceneNode(m, 0, 0, 1, 1, Isometry3.identity(), Ball(2.0)),
             SceneNode(m2, 0, 0, 1, 1, Isometry3.identity(), Ball(3.0))]
    flat = Scene(nodes, [Light((0, 0, 0), 0, 1, (1, 1, 1))], upload=False).flat
    assert flat.desc.n_materials == 2 and flat.desc.n_textures == 1 and flat.desc.n_texels == 4
    assert flat.node_rows[0].material == flat.node_rows[1].material != flat.node_rows[2].material


def test_user_defined_material_cannot_cross_the_abi():
    class Mine:
        pass
    with pytest.raises(TypeError):
        SceneNode(Mine(), 0, 0, 1, 1, Isometry3.identity(), Ball(1.0))


def test_trimesh_validation():
    with pytest.raises(ValueError):
        TriMesh(np.zeros((3, 3), np.float32), np.array([[0, 1, 3]], np.uint32))
    with pytest.raises(ValueError):
        TriMesh(np.zeros((3, 3), np.float32), np.array([[0, 1, 2]], np.uint32), np.zeros((2, 2), np.float32))


@pytest.mark.parametrize("w,h,world", [(64, 48, 1), (50, 37, 2), (33, 16, 3), (1920, 1080, 8)])
def test_tile_pack_untile_roundtrip(w, h, world):
    rng = np.random.default_rng(0)
    img = rng.uniform(size=(h, w, 3)).astype(np.float32)
    packed = np.stack([dist.pack_tiles_host(img, r, world) for r in range(world)])
    assert packed.shape[1] == dist.tiles_per_rank(w, h, world)
    np.testing.assert_array_equal(dist.untile_host(packed, world, w, h), img)
    owners = sorted(t for r in range(world) for t in dist.local_tiles(w, h, r, world))
    assert owners == list(range(dist.tile_count(w, h)))            # every tile owned exactly once


def test_baseline_configs_build():
    for name in ("C1", "C2"):
        scene, cam, cfg = configs.build_flat(name, globe_size=(64, 32))
        assert cfg["width"] > 0 and len(scene.nodes) in (3, 5)
    scene, cam, cfg = configs.build_flat("C3", target_tris=30000, lod=4)
    assert scene.flat.n_triangles == 30000 and len(scene.nodes) >= 20
    assert sum(1 for n in scene.nodes if n.material.alpha is not None) == 2   # two alpha-mapped groups
    assert cam.eye == (-250.0, 50.0, 0.0)
