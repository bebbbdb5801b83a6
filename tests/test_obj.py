"""OBJ / MTL ingest twin (src/obj.rs, src/mtl.rs) -> flattened scene (§8f rank 3)."""
import numpy as np
import pytest

import oracle_lib as O
from nrays_b200 import assets, configs, make_camera, obj
from nrays_b200.loader3d import AssetResolver, SceneFileError, load_scene

OBJ = """
# two groups, quad + pentagon, relative indices, one material switch
mtllib m.mtl
v 0 0 0
v 1 0 0
v 1 1 0
v 0 1 0
v 2 0 0
v 2 1 0
v 3 0.5 0
vt 0 0
vt 1 0
vt 1 1
vt 0 1
g left
usemtl red
f 1/1 2/2 3/3 4/4
g right
usemtl blue
f 2/2 5/1 6/4 -5/3
f -6/2 5/1 7/2 6/4 3/3
"""
MTL = "newmtl red\nKd 1 0 0\nnewmtl blue\nKd 0 0 1\nd 0.5\n"


def test_groups_fan_triangulation_and_dedup():
    od = obj.parse(OBJ, lambda name: MTL, "base")
    names = [g[0] for g in od.groups]
    # `g right` inherits the current material (red, src/obj.rs:95), so the following `usemtl blue` is a SECOND
    # usemtl for that group and opens an auto-generated group (src/obj.rs:149-160); empty groups are dropped
    assert names == ["base/left", "auto_generated_group_/2blue"]
    left, right = od.groups[0][1], od.groups[1][1]
    assert left.shape == (2, 3) and right.shape == (2 + 3, 3)   # quad -> 2, pentagon -> 3 triangles
    assert od.groups[0][2].name == "red" and od.groups[1][2].name == "blue" and od.groups[1][2].alpha == 0.5
    # quad fan pivots on its first vertex: (0,1,2) (0,2,3)
    c = od.coords
    np.testing.assert_allclose(c[left[0]], [[0, 0, 0], [1, 0, 0], [1, 1, 0]])
    np.testing.assert_allclose(c[left[1]], [[0, 0, 0], [1, 1, 0], [0, 1, 0]])
    # relative index -5 (of 7 vertices) = vertex 3 -> (1,1,0)
    np.testing.assert_allclose(c[right[1]][2], [1, 1, 0])
    # pentagon: the reference pivots on g[len - i]: third triangle starts at the face's THIRD emitted entry
    pent = right[2:]
    np.testing.assert_allclose(c[pent[0]], [[1, 0, 0], [2, 0, 0], [3, 0.5, 0]])
    np.testing.assert_allclose(c[pent[1]][0], [1, 0, 0])        # quad-like second triangle: pivot = first vertex
    np.testing.assert_allclose(c[pent[2]][0], [3, 0.5, 0])      # fifth vertex: pivot = g[len-4] (mis-triangulation kept)
    # de-duplication on the (v, vt, vn) triple: vertex 2 with vt 2 is shared by both groups
    assert len(od.coords) == len({tuple(r) for r in np.concatenate([od.coords, od.uvs], 1)})


def test_missing_vt_drops_all_uvs_and_second_usemtl_makes_a_group():
    text = "v 0 0 0\nv 1 0 0\nv 0 1 0\nvt 0 0\nmtllib m.mtl\ng a\nusemtl red\nf 1/1 2/1 3/1\nusemtl blue\nf 1 2 3\n"
    od = obj.parse(text, lambda name: MTL, "")
    assert od.uvs is None                                        # one face vertex without vt: uvs dropped file-wide
    assert [g[0] for g in od.groups] == ["/a", "auto_generated_group_/1blue"]
    assert od.groups[1][2].name == "blue"


def test_bad_input_raises_where_the_reference_panics():
    with pytest.raises(SceneFileError):
        obj.parse("v 0 0\n")
    with pytest.raises(SceneFileError):
        obj.parse("v 0 0 0\nf 1 2 x\n")
    with pytest.raises(SceneFileError):
        obj.parse("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 9\n")


def test_standin_through_obj_text_gives_the_same_render():
    """Stand-in mesh -> OBJ/MTL text -> parser twin -> scene == stand-in handed over directly (oracle images equal)."""
    od, tex = assets.sponza_standin(seed=0, target_tris=9000, lod=8)
    obj_text, mtl_text = obj.write_obj(od, "sponza.mtl")
    r = AssetResolver()
    r.files["media/crytek-sponza/sponza.obj"] = obj_text
    r.files["media/crytek-sponza/sponza.mtl"] = mtl_text
    for k, v in tex.items():
        r.textures["media/crytek-sponza/" + k] = v
    scene_txt, cams = load_scene(configs.sponza_text(), r, upload=False)
    scene_dir, _ = load_scene(configs.sponza_text(), configs.sponza_resolver(target_tris=9000, lod=8), upload=False)
    assert scene_txt.flat.n_triangles == scene_dir.flat.n_triangles == 9000
    assert len(scene_txt.nodes) == len(scene_dir.nodes)
    w, h = 64, 36
    cam = make_camera(w, h, 1, 0.0, cams[0].eye, cams[0].projection((w, h)))
    a, sa = O.OracleScene(scene_txt.flat, 64).render(cam)
    b, sb = O.OracleScene(scene_dir.flat, 64).render(cam)
    np.testing.assert_allclose(a, b, atol=1e-6)
    assert sa.rays_total == sb.rays_total
