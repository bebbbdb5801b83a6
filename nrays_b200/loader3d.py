"""Twin of examples/loader3d.rs: `.scene` parser, scene assembly and camera setup.

The reference front end is the only caller of scene::render (examples/loader3d.rs:86-93); this
module exists so the same `.scene` files drive the B200 path.  Grammar: SURVEY.md Appendix C
(examples/loader3d.rs:214-346); defaults: aa=(1,0) :426, radius=0/nsample=1 :452-453,
refl=(0,0) :555-558, refr=1.0 :562; `angle` is an axis-angle vector in degrees :546-552.

Assets that the reference reads from disk (`media/...`, git-ignored and absent from the tree) are
obtained through an AssetResolver, which serves registered synthetic stand-ins first and falls
back to the file system.
"""
import math
import os

import numpy as np

from .scene import (Ball, Capsule, Cone, Cuboid, Cylinder, Interpolation, Isometry3, Light, NormalMaterial, Overflow,
                    PhongMaterial, Plane, Scene, SceneNode, Texture2d, TriMesh, UVMaterial, camera_projection, render)


class SceneFileError(Exception):
    """Raised where the reference panics (`error(line, ..)`, examples/loader3d.rs:206-208)."""


class MtlMaterial:
    """MtlMaterial::new_default — src/mtl.rs:147-162."""

    def __init__(self, name):
        self.name = name
        self.shininess = 60.0
        self.alpha = 1.0
        self.ambiant_texture = None
        self.diffuse_texture = None
        self.specular_texture = None
        self.opacity_map = None
        self.ambiant = (1.0, 1.0, 1.0)
        self.diffuse = (1.0, 1.0, 1.0)
        self.specular = (1.0, 1.0, 1.0)


def parse_mtl(string):
    """mtl::parse — src/mtl.rs:29-89."""
    res = []
    cur = MtlMaterial("")
    for l, line in enumerate(string.splitlines()):
        words = line.split()
        if not words or words[0].startswith("#") or len(words) < 2:
            continue
        tag, rest = words[0], words[1:]

        def color():
            if len(rest) < 3:
                raise SceneFileError("At line %d: 3 components were expected, found %d." % (l, len(rest)))
            return tuple(float(np.float32(x)) for x in rest[:3])

        if tag == "newmtl":
            if cur.name:
                res.append(cur)
            cur = MtlMaterial(" ".join(rest))
        elif tag == "Ka":
            cur.ambiant = color()
        elif tag == "Kd":
            cur.diffuse = color()
        elif tag == "Ks":
            cur.specular = color()
        elif tag == "Ns":
            cur.shininess = float(np.float32(rest[0]))
        elif tag == "d":
            cur.alpha = float(np.float32(rest[0]))
        elif tag == "map_Ka":
            cur.ambiant_texture = " ".join(rest)
        elif tag == "map_Kd":
            cur.diffuse_texture = " ".join(rest)
        elif tag == "map_Ks":
            cur.specular_texture = " ".join(rest)
        elif tag in ("map_d", "map_opacity"):
            cur.opacity_map = " ".join(rest)
    if cur.name:
        res.append(cur)
    return res


class ObjData:
    """What obj::parse_file yields for one file (src/obj.rs:62-120, 327-397): one shared, deduplicated
    vertex/uv array and per-group face lists with their MTL material."""

    def __init__(self, coords, uvs, groups):
        self.coords = np.ascontiguousarray(coords, dtype=np.float32).reshape(-1, 3)
        self.uvs = None if uvs is None else np.ascontiguousarray(uvs, dtype=np.float32).reshape(-1, 2)
        self.groups = groups  # list of (name, faces (F,3) uint32, MtlMaterial | None)


class AssetResolver:
    """Serves textures / OBJ files by the path written in the scene or MTL file."""

    def __init__(self, base_dir="."):
        self.base_dir = base_dir
        self.textures = {}  # path -> (H,W[,depth]) uint8 array (pre-flip, as a PNG decoder would return)
        self.objs = {}      # path -> ObjData
        self.files = {}     # path -> text (mtl files)
        self._tex_cache = {}

    def read_text(self, path):
        if path in self.files:
            return self.files[path]
        try:
            with open(os.path.join(self.base_dir, path), "r") as f:
                return f.read()
        except OSError as e:
            raise OSError("cannot read %s: %s" % (path, e))

    def texture(self, path, opacity):
        """Texture2d::from_png(path, opacity, Bilinear, Wrap) with the per-path cache of
        src/texture2d.rs:28-48 (one ImageData per (path, opacity))."""
        key = (path, bool(opacity))
        if key in self._tex_cache:
            data = self._tex_cache[key]
            return Texture2d(data, Interpolation.Bilinear, Overflow.Wrap)
        if path in self.textures:
            t = Texture2d.from_array(self.textures[path], opacity, Interpolation.Bilinear, Overflow.Wrap)
        else:
            t = Texture2d.from_png(os.path.join(self.base_dir, path), opacity, Interpolation.Bilinear, Overflow.Wrap)
        if t is None:
            raise SceneFileError("Image not found: %s" % path)
        self._tex_cache[key] = t.data
        return t

    def obj(self, objpath, mtldir):
        """obj::parse_file(objpath, mtldir, "") — loader3d.rs:662.  Registered stand-ins first, then OBJ text
        registered under `files`, then the file system."""
        if objpath in self.objs:
            return self.objs[objpath]
        from . import obj as objmod

        if objpath in self.files:
            return objmod.parse(self.files[objpath], lambda name: self.read_text(os.path.join(mtldir, name)), "")
        full = os.path.join(self.base_dir, objpath)
        if not os.path.exists(full):
            raise SceneFileError("OBJ file %s not found (scenes/media/ is not part of the reference tree)" % objpath)
        return objmod.parse_file(full, os.path.join(self.base_dir, mtldir), "")


class Camera:
    def __init__(self, eye, at, fovy, resolution, aa, output):
        if not aa[0] >= 1.0:
            raise SceneFileError("The number of ray per pixel must be at least 1.0")  # loader3d.rs:146-149
        self.eye, self.at, self.fovy = tuple(eye), tuple(at), float(fovy)
        self.resolution = (float(resolution[0]), float(resolution[1]))
        self.aa = (float(aa[0]), float(aa[1]))
        self.output = output

    def projection(self, resolution=None):
        w, h = resolution if resolution is not None else self.resolution
        return camera_projection(self.eye, self.at, self.fovy, w, h)


def _floats(l, words, n):
    if len(words) < n:
        raise SceneFileError("At line %d: %d components were expected, found %d." % (l, n, len(words)))
    try:
        return [float(w) for w in words[:n]]
    except ValueError as e:
        raise SceneFileError("At line %d: failed to parse as a f64: %s" % (l, e))


def parse(string, resolver=None):
    """parse — examples/loader3d.rs:214-346.  Returns (lights, nodes, cameras)."""
    resolver = resolver or AssetResolver()
    nodes, lights, cameras = [], [], []
    # built-in materials — loader3d.rs:226-249
    mtllib = {
        "normals": (1.0, NormalMaterial()),
        "uvs": (1.0, UVMaterial()),
        "default": (1.0, PhongMaterial((0.1, 0.1, 0.1), (1.0, 1.0, 1.0), (1.0, 1.0, 1.0), None, None, 100.0)),
    }
    mode = None
    props = {"superbloc": 0, "geom": []}

    def register():
        if mode == "light":
            _register_light(props, lights)
        elif mode == "geometry":
            _register_geometry(props, mtllib, nodes, resolver)
        elif mode == "camera":
            _register_camera(props, cameras)

    for l, line in enumerate(string.splitlines()):
        words = line.split()
        if not words or words[0].startswith("#"):
            continue
        tag, rest = words[0], words[1:]
        if tag == "mtllib":
            _register_mtllib(" ".join(rest), mtllib, resolver)
        elif tag in ("light", "geometry", "camera"):
            register()
            props = {"superbloc": l, "geom": []}
            mode = tag
        elif tag in ("color", "angle", "pos", "eye", "at"):
            props[tag] = _floats(l, rest, 3)
        elif tag in ("material", "output"):
            props[tag] = " ".join(rest)
        elif tag in ("fovy", "refr", "radius", "nsample"):
            props[tag] = _floats(l, rest, 1)[0]
        elif tag in ("resolution", "refl", "aa"):
            props[tag] = _floats(l, rest, 2)
        elif tag == "ball":
            props["geom"].append(("ball", _floats(l, rest, 1)))
        elif tag == "plane":
            props["geom"].append(("plane", _floats(l, rest, 3)))
        elif tag == "box":
            props["geom"].append(("box", _floats(l, rest, 3)))
        elif tag in ("cylinder", "capsule", "cone"):
            props["geom"].append((tag, _floats(l, rest, 2)))
        elif tag == "obj":
            if len(rest) < 2:
                raise SceneFileError("At line %d: 2 paths were expected, found %d." % (l, len(rest)))
            props["geom"].append(("obj", rest[:2]))
        elif tag == "solid":
            props["solid"] = True
        # unknown lines are ignored with a warning in the reference (:327-329)
    register()
    return lights, nodes, cameras


def _need(props, key, what):
    if key not in props:
        raise SceneFileError("At line %d: missing attribute: %s" % (props["superbloc"], what))


def _register_camera(props, cameras):  # loader3d.rs:408-436
    for k, what in (("output", "output <filename>"), ("resolution", "resolution <x> <y>"), ("eye", "eye <x> <y> <z>"),
                    ("at", "at <x> <y> <z>"), ("fovy", "fovy <value>")):
        _need(props, k, what)
    aa = props.get("aa", [1.0, 0.0])
    cameras.append(Camera(props["eye"], props["at"], props["fovy"], props["resolution"], aa, props["output"]))


def _register_light(props, lights):  # loader3d.rs:438-459
    _need(props, "pos", "pos <x> <y> <z>")
    _need(props, "color", "color <r> <g> <b>")
    radius = props.get("radius", 0.0)
    nsample = props.get("nsample", 1.0)
    color = [float(np.float32(c)) for c in props["color"]]
    lights.append(Light(props["pos"], radius, int(nsample), color))


def _phong_from_mtl(m, resolver, prefix=None):
    def path(p):
        return p if prefix is None else os.path.join(prefix, p)

    t = resolver.texture(path(m.diffuse_texture), False) if m.diffuse_texture else None
    a = resolver.texture(path(m.opacity_map), True) if m.opacity_map else None
    return PhongMaterial(m.ambiant, m.diffuse, m.specular, t, a, m.shininess)


def _register_mtllib(path, mtllib, resolver):  # loader3d.rs:461-503
    for m in parse_mtl(resolver.read_text(path)):
        mtllib[m.name] = (m.alpha, _phong_from_mtl(m, resolver))


def _register_geometry(props, mtllib, nodes, resolver):  # loader3d.rs:505-792
    _need(props, "pos", "pos <x> <y> <z>")
    _need(props, "angle", "color <r> <g> <b>")  # sic: the reference's message
    if not props["geom"]:
        raise SceneFileError("At line %d: missing attribute: <geom_type> <geom parameters>]" % props["superbloc"])
    _need(props, "material", "material <material_name>")
    solid = bool(props.get("solid", False))
    mname = props["material"]
    special = mname in ("uvs", "normals")
    if mname not in mtllib:
        raise SceneFileError("Attempted to use an unknown material: %s" % mname)
    alpha, material = mtllib[mname]
    alpha = float(np.float32(alpha))
    angle = [math.radians(a) for a in props["angle"]]
    transform = Isometry3.new(props["pos"], angle)
    refl = props.get("refl", [0.0, 0.0])
    refl_m, refl_a = float(np.float32(refl[0])), float(np.float32(refl[1]))
    refr_c = float(props.get("refr", 1.0))

    kind, p = props["geom"][0]  # only the first shape of a block is used (F11, loader3d.rs:593)

    def push(geom, mat=material, a=alpha):
        nodes.append(SceneNode(mat, refl_m, refl_a, a, refr_c, transform, geom, None, solid))

    if kind == "ball":
        push(Ball(p[0]))
    elif kind == "box":
        push(Cuboid(p))
    elif kind == "cylinder":
        push(Cylinder(p[0], p[1]))
    elif kind == "capsule":
        push(Capsule(p[0], p[1]))
    elif kind == "cone":
        push(Cone(p[0], p[1]))
    elif kind == "plane":
        push(Plane(p))
    elif kind == "obj":
        od = resolver.obj(p[0], p[1])
        # vertices / 4 in f64 then (here) back to f32 — exact: loader3d.rs:665-670
        coords = (od.coords.astype(np.float64) / 4.0).astype(np.float32)
        uvs = od.uvs if od.uvs is not None else np.zeros((len(coords), 2), np.float32)
        for _name, faces, mat in od.groups:
            if len(faces) == 0:
                continue
            mesh = TriMesh(coords, faces, uvs)
            if mat is not None:
                color = _phong_from_mtl(mat, resolver, prefix=p[1])
                push(mesh, material if special else color, float(np.float32(mat.alpha) * np.float32(alpha)))
            else:
                push(mesh)


def load_scene(text, resolver=None, device=0, upload=True):
    """main() up to Scene::new — loader3d.rs:57-61 (background = (1,1,1))."""
    lights, nodes, cameras = parse(text, resolver)
    scene = Scene(nodes, lights, (1.0, 1.0, 1.0), device=device, upload=upload)
    return scene, cameras


def render_camera(scene, camera, resolution=None, aa=None, seed=0, return_stats=False):
    """The per-camera body of main() — loader3d.rs:67-93, with optional resolution / aa overrides
    (the BASELINE configs override both; the scene files ship other values)."""
    res = resolution if resolution is not None else camera.resolution
    aa = aa if aa is not None else camera.aa
    proj = camera.projection(res)
    return render(scene, res, int(aa[0]), aa[1], camera.eye, proj, seed=seed, return_stats=return_stats)
