"""The C++ host mirror (nrays_b200/host/nrays.hpp + loader3d.cpp, twin of examples/loader3d.rs) against the
Python host on the same scene files: identical flattened tables (CPU) and identical images (GPU)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from nrays_b200 import _abi as A, _lib, assets, obj
from nrays_b200.loader3d import AssetResolver, load_scene

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LOADER = os.path.join(ROOT, "nrays_b200", "host", "loader3d")

SCENE = """mtllib mats.mtl
camera
  output out.png
  resolution 96 64
  eye 30 60 -260
  at 0 40 0
  fovy 45
light
  pos 20 120 -100
  color 1 0.9 0.8
  radius 3
  nsample 4
geometry
  obj mesh.obj assets
  pos 0 0 0
  angle 0 10 0
  material default
geometry
  ball 25
  pos -60 30 -120
  angle 0 0 0
  material tex
  refl 0.3 0.4
geometry
  box 20 20 20
  pos 70 25 -130
  angle 0 30 0
  material glass
  refr 1.3
geometry
  cone 25 15
  pos 0 25 -150
  angle 0 0 0
  material normals
geometry
  plane 0 1 0
  pos 0 -2.5 0
  angle 0 0 0
  material default
  refl 0.2 0.5
"""
MATS = "newmtl tex\nKa 0.3 0.3 0.3\nKd 0.9 0.9 0.9\nmap_Kd tex.png\nnewmtl glass\nd 0.4\nKa 0.1 0.2 0.3\nKd 0.4 0.6 0.9\nNs 80\n"


@pytest.fixture(scope="module")
def scene_dir(tmp_path_factory):
    from PIL import Image

    d = tmp_path_factory.mktemp("cpp_host")
    (d / "assets" / "textures").mkdir(parents=True)
    od, tex = assets.sponza_standin(seed=0, target_tris=9000, lod=8)
    obj_text, mtl_text = obj.write_obj(od, "mesh.mtl")
    (d / "mesh.obj").write_text(obj_text)
    (d / "assets" / "mesh.mtl").write_text(mtl_text)
    for k, v in tex.items():
        Image.fromarray(v).save(str(d / "assets" / k))
    Image.fromarray(assets.globe_texture(64, 32)).save(str(d / "tex.png"))
    (d / "mats.mtl").write_text(MATS)
    (d / "t.scene").write_text(SCENE)
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "nrays_b200", "host")], stdout=subprocess.DEVNULL)
    return str(d)


def test_cpp_and_python_hosts_flatten_to_the_same_tables(scene_dir):
    out = subprocess.check_output([LOADER, "t.scene", "--validate"], cwd=scene_dir, text=True)
    m = re.search(r"VALID (.*)", out)
    assert m, out
    cpp = {k: int(v) for k, v in (kv.split("=") for kv in m.group(1).split())}
    scene, cams = load_scene(SCENE, AssetResolver(base_dir=scene_dir), upload=False)
    d = scene.flat.desc
    info = A.NrbBuildInfo()
    assert _lib.load().nrb_scene_validate(C.byref(d), C.byref(info)) == 0
    py = dict(nodes=d.n_nodes, lights=d.n_lights, materials=d.n_materials, textures=d.n_textures, texels=d.n_texels,
              vertices=d.n_vertices, indices=d.n_indices, bvh_nodes=info.bvh_nodes, triangles=info.triangles,
              shapes=info.shapes, planes=info.planes, candidates=info.transparent_candidates, depth=info.max_depth)
    assert cpp == py
    assert py["triangles"] == 9000 and py["planes"] == 1 and py["textures"] >= 3


def test_cpp_host_reports_errors_like_the_reference_panics(scene_dir, tmp_path):
    bad = tmp_path / "bad.scene"
    bad.write_text("geometry\n ball 1\n pos 0 0 0\n angle 0 0 0\n material nope\n")
    r = subprocess.run([LOADER, str(bad), "--validate"], capture_output=True, text=True)
    assert r.returncode == 101 and "unknown material" in r.stderr
    r = subprocess.run([LOADER, "/nonexistent.scene"], capture_output=True, text=True)
    assert r.returncode == 101 and "Unable to find the file" in r.stderr
    if _lib.load().nrb_device_count() == 0:   # no CPU fallback behind the C++ host either
        r = subprocess.run([LOADER, "t.scene"], cwd=scene_dir, capture_output=True, text=True)
        assert r.returncode == 101 and "no CUDA device" in r.stderr


@pytest.mark.gpu
def test_cpp_host_renders_the_same_image(gpu, scene_dir):
    from PIL import Image

    from nrays_b200.loader3d import render_camera

    subprocess.check_call([LOADER, "t.scene", "--aa", "2", "1.0", "--seed", "5"], cwd=scene_dir, stdout=subprocess.DEVNULL)
    cpp = np.asarray(Image.open(os.path.join(scene_dir, "out.png"))).astype(int)
    scene, cams = load_scene(SCENE, AssetResolver(base_dir=scene_dir))
    img = render_camera(scene, cams[0], aa=(2, 1.0), seed=5)
    py = img.to_rgb8().astype(int)
    assert cpp.shape == py.shape == (64, 96, 3)
    diff = np.abs(cpp - py)
    assert diff.max() <= 1 and (diff > 0).mean() < 0.01   # same library, same tables: only float-atomic order differs
    scene.close()
