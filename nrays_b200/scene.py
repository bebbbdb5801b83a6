"""Host-side mirror of the reference's scene-building API, above the C-ABI.

The reference host is Rust (no toolchain in this image), so this module keeps the same type and
argument names as the Rust crate and flattens them into the POD tables of include/nrays_b200.h:

    Light::new(pos, radius, nsample, color)                          src/light.rs:16-23
    PhongMaterial::new(ambiant, diffuse, specular, texture, alpha, shininess)
                                                                     src/phong_material.rs:19-35
    NormalMaterial::new(), UVMaterial::new()                         src/normal_material.rs, src/uv_material.rs
    Texture2d::new(data, interpolation, overflow), ImageData::new    src/texture2d.rs:10-76
    SceneNode::new(material, refl_mix, refl_atenuation, alpha, refr_coeff, transform,
                   geometry, nmap, solid)                            src/scene_node.rs:22-47
    Scene::new(nodes, lights, background)                            src/scene.rs:119-133
    render(scene, resolution, ray_per_pixel, window_width, camera_eye, projection) -> Image
                                                                     src/scene.rs:29-116
    Ball / Cuboid / Cylinder / Capsule / Cone / Plane / TriMesh      ncollide3d shapes, loader3d.rs:593-695

Nothing here computes pixels: `render` calls nrb_render through ctypes and raises if the CUDA
library is missing or reports an error (the reference panics in the same places).
"""
import ctypes as C
import math

import numpy as np

from . import _abi as A


# ---------------------------------------------------------------------------------------------
# math helpers (nalgebra 0.15 equivalents used by examples/loader3d.rs:68-79, 546-552)
# ---------------------------------------------------------------------------------------------
class Isometry3:
    """nalgebra Isometry3<f64>: rotation (stored as a 3x3 matrix) + translation."""

    def __init__(self, rot=None, trans=None):
        self.rot = np.eye(3) if rot is None else np.asarray(rot, dtype=np.float64).reshape(3, 3)
        self.trans = np.zeros(3) if trans is None else np.asarray(trans, dtype=np.float64).reshape(3)

    @staticmethod
    def identity():
        return Isometry3()

    @staticmethod
    def new(translation, axisangle):
        """Isometry3::new(translation, axisangle): rotation = exp map of the axis-angle vector."""
        w = np.asarray(axisangle, dtype=np.float64)
        ang = float(np.linalg.norm(w))
        if ang == 0.0:
            R = np.eye(3)
        else:
            k = w / ang
            K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
            R = np.eye(3) + math.sin(ang) * K + (1.0 - math.cos(ang)) * (K @ K)
        return Isometry3(R, translation)

    @staticmethod
    def look_at_rh(eye, target, up):
        """Isometry3::look_at_rh — SURVEY B.9: rows (right, up', -forward), translation R*(-eye)."""
        eye = np.asarray(eye, dtype=np.float64)
        f = np.asarray(target, dtype=np.float64) - eye
        f = f / np.linalg.norm(f)
        r = np.cross(f, np.asarray(up, dtype=np.float64))
        r = r / np.linalg.norm(r)
        u = np.cross(r, f)
        R = np.stack([r, u, -f])
        return Isometry3(R, R @ (-eye))

    def to_homogeneous(self):
        M = np.eye(4)
        M[:3, :3] = self.rot
        M[:3, 3] = self.trans
        return M


def perspective3(aspect, fovy, znear, zfar):
    """Perspective3::new(aspect, fovy, znear, zfar).to_homogeneous() — SURVEY B.9 (OpenGL matrix)."""
    t = math.tan(fovy / 2.0)
    P = np.zeros((4, 4))
    P[0, 0] = 1.0 / (aspect * t)
    P[1, 1] = 1.0 / t
    P[2, 2] = (zfar + znear) / (znear - zfar)
    P[2, 3] = 2.0 * zfar * znear / (znear - zfar)
    P[3, 2] = -1.0
    return P


def camera_projection(eye, at, fovy_deg, width, height):
    """The matrix examples/loader3d.rs:68-79 hands to render: (perspective * view)^-1."""
    persp = perspective3(float(width) / float(height), math.radians(fovy_deg), 1.0, 100000.0)
    view = Isometry3.look_at_rh(eye, at, (0.0, 1.0, 0.0)).to_homogeneous()
    return np.linalg.inv(persp @ view)


# ---------------------------------------------------------------------------------------------
# shapes (ncollide3d::shape::*)
# ---------------------------------------------------------------------------------------------
class Ball:
    kind = A.NRB_SHAPE_BALL

    def __init__(self, radius):
        self.param = (float(radius), 0.0, 0.0)


class Cuboid:
    kind = A.NRB_SHAPE_CUBOID

    def __init__(self, half_extents):
        self.param = tuple(float(x) for x in half_extents)


class Cylinder:
    kind = A.NRB_SHAPE_CYLINDER

    def __init__(self, half_height, radius):
        self.param = (float(half_height), float(radius), 0.0)


class Capsule:
    kind = A.NRB_SHAPE_CAPSULE

    def __init__(self, half_height, radius):
        self.param = (float(half_height), float(radius), 0.0)


class Cone:
    kind = A.NRB_SHAPE_CONE

    def __init__(self, half_height, radius):
        self.param = (float(half_height), float(radius), 0.0)


class Plane:
    kind = A.NRB_SHAPE_PLANE

    def __init__(self, normal):
        n = np.asarray(normal, dtype=np.float64)
        n = n / np.linalg.norm(n)  # Unit::new_normalize, loader3d.rs:656
        self.param = tuple(float(x) for x in n)


class TriMesh:
    """TriMesh::new(coords, faces, uvs) — loader3d.rs:695.  `coords` (V,3), `faces` (F,3), `uvs` (V,2)|None.

    Several meshes may share one coords/uvs array object (the reference clones the whole vertex
    array per OBJ group, loader3d.rs:690-695); flatten() stores a shared array once.
    """

    kind = A.NRB_SHAPE_TRIMESH
    param = (0.0, 0.0, 0.0)

    def __init__(self, coords, faces, uvs=None):
        self.coords = np.ascontiguousarray(coords, dtype=np.float32).reshape(-1, 3)
        self.faces = np.ascontiguousarray(faces, dtype=np.uint32).reshape(-1, 3)
        self.uvs = None if uvs is None else np.ascontiguousarray(uvs, dtype=np.float32).reshape(-1, 2)
        if self.faces.size and int(self.faces.max()) >= len(self.coords):
            raise ValueError("TriMesh face index out of range")
        if self.uvs is not None and len(self.uvs) != len(self.coords):
            raise ValueError("TriMesh uvs must be per-vertex")
        # meshes built on the same vertex buffer share one copy in the flattened tables
        self._coords_key = (self.coords.__array_interface__["data"][0], len(self.coords))


# ---------------------------------------------------------------------------------------------
# textures / materials / lights
# ---------------------------------------------------------------------------------------------
class Interpolation:
    Bilinear = A.NRB_INTERP_BILINEAR
    Nearest = A.NRB_INTERP_NEAREST


class Overflow:
    ClampToEdges = A.NRB_OVERFLOW_CLAMP
    Wrap = A.NRB_OVERFLOW_WRAP


class ImageData:
    """ImageData::new(pixels, dims) — src/texture2d.rs:10-26. pixels: (H*W,4) f32 RGBA, row-major y*W+x."""

    def __init__(self, pixels, dims):
        w, h = int(dims[0]), int(dims[1])
        self.pixels = np.ascontiguousarray(pixels, dtype=np.float32).reshape(-1, 4)
        assert len(self.pixels) == w * h
        assert w >= 1 and h >= 1
        self.dims = (w, h)


class Texture2d:
    def __init__(self, data, interpolation, overflow):
        self.data = data
        self.interpol = int(interpolation)
        self.overflow = int(overflow)

    @staticmethod
    def new(data, interpolation, overflow):
        return Texture2d(data, interpolation, overflow)

    @staticmethod
    def from_array(img_u8, opacity, interpolation, overflow):
        """The decode conventions of Texture2d::from_png (src/texture2d.rs:96-173) applied to an
        already-decoded (H,W[,depth]) uint8 array: y flip, then channel expansion by depth."""
        img = np.asarray(img_u8)
        if img.ndim == 2:
            img = img[:, :, None]
        img = img[::-1].astype(np.float32) / 255.0  # flip the y axis (:99-107)
        h, w, depth = img.shape
        out = np.ones((h, w, 4), dtype=np.float32)
        if depth == 1:
            g = img[:, :, 0]
            if opacity:
                out[:, :, 3] = g
            else:
                out[:, :, 0] = out[:, :, 1] = out[:, :, 2] = g
        elif depth == 2:
            rg = img[:, :, 0] * img[:, :, 1]
            if opacity:
                out[:, :, 3] = img[:, :, 1] * img[:, :, 0]
            else:
                out[:, :, 0] = out[:, :, 1] = out[:, :, 2] = rg
        elif depth == 3:
            if opacity:
                out[:, :, 3] = img[:, :, 0]  # red channel: texture2d.rs:140
            else:
                out[:, :, :3] = img
        elif depth == 4:
            if opacity:
                out[:, :, 3] = img[:, :, 3]
            else:
                out[:, :, :3] = img[:, :, :3]  # alpha dropped: texture2d.rs:163
        else:
            raise ValueError("Image depth %d not suported." % depth)
        return Texture2d(ImageData(out.reshape(-1, 4), (w, h)), interpolation, overflow)

    @staticmethod
    def from_png(path, opacity, interpolation, overflow):
        from PIL import Image as PILImage

        try:
            im = PILImage.open(path)
            im.load()
        except Exception:
            return None
        if im.mode not in ("L", "LA", "RGB", "RGBA"):
            im = im.convert("RGBA" if "A" in im.getbands() else "RGB")
        return Texture2d.from_array(np.asarray(im), opacity, interpolation, overflow)


class Material:
    """trait Material (src/material.rs:6-17).  Only the three reference impls can cross the C-ABI."""

    kind = None


class PhongMaterial(Material):
    kind = A.NRB_MAT_PHONG

    def __init__(self, ambiant_color, diffuse_color, specular_color, texture, alpha, shininess):
        self.ambiant_color = tuple(float(x) for x in ambiant_color)
        self.diffuse_color = tuple(float(x) for x in diffuse_color)
        self.specular_color = tuple(float(x) for x in specular_color)
        self.texture = texture
        self.alpha = alpha
        self.shininess = float(shininess)


class NormalMaterial(Material):
    kind = A.NRB_MAT_NORMAL


class UVMaterial(Material):
    kind = A.NRB_MAT_UV


class Light:
    def __init__(self, pos, radius, nsample, color):
        self.pos = tuple(float(x) for x in pos)
        self.radius = float(radius)
        # ((nsample as f32).sqrt()) as usize — src/light.rs:20
        self.racsample = int(np.sqrt(np.float32(int(nsample))))
        self.color = tuple(float(x) for x in color)

    new = None  # set below


Light.new = staticmethod(lambda pos, radius, nsample, color: Light(pos, radius, nsample, color))


class SceneNode:
    def __init__(self, material, refl_mix, refl_atenuation, alpha, refr_coeff, transform, geometry, nmap=None,
                 solid=False):
        if not isinstance(material, Material) or material.kind is None:
            raise TypeError("only PhongMaterial / NormalMaterial / UVMaterial can be flattened across the C-ABI")
        self.material = material
        self.refl_mix = float(refl_mix)
        self.refl_atenuation = float(refl_atenuation)
        self.alpha = float(alpha)
        self.refr_coeff = float(refr_coeff)
        self.transform = transform
        self.geometry = geometry
        self.nmap = nmap
        self.solid = bool(solid)

    new = None


SceneNode.new = staticmethod(lambda *a, **k: SceneNode(*a, **k))


# ---------------------------------------------------------------------------------------------
# flattening
# ---------------------------------------------------------------------------------------------
class FlatScene:
    """The POD tables of NrbSceneDesc plus the numpy buffers that back its pointers."""

    def __init__(self, nodes, lights, background):
        self._keep = []
        tex_index = {}
        tex_rows = []
        texel_chunks = []
        texel_off = 0
        data_off = {}

        def tex_id(t):
            nonlocal texel_off
            if t is None:
                return -1
            key = (id(t.data), t.interpol, t.overflow)
            if key in tex_index:
                return tex_index[key]
            if id(t.data) not in data_off:
                data_off[id(t.data)] = texel_off
                texel_chunks.append(t.data.pixels)
                texel_off += len(t.data.pixels)
            row = A.NrbTextureDesc(t.data.dims[0], t.data.dims[1], t.interpol, t.overflow, data_off[id(t.data)])
            tex_index[key] = len(tex_rows)
            tex_rows.append(row)
            return tex_index[key]

        mat_index = {}
        mat_rows = []

        def mat_id(m):
            if id(m) in mat_index:
                return mat_index[id(m)]
            row = A.NrbMaterialDesc()
            row.kind = m.kind
            row.texture = -1
            row.alpha_texture = -1
            if m.kind == A.NRB_MAT_PHONG:
                row.ambient = (C.c_float * 3)(*m.ambiant_color)
                row.diffuse = (C.c_float * 3)(*m.diffuse_color)
                row.specular = (C.c_float * 3)(*m.specular_color)
                row.shininess = m.shininess
                row.texture = tex_id(m.texture)
                row.alpha_texture = tex_id(m.alpha)
            mat_index[id(m)] = len(mat_rows)
            mat_rows.append(row)
            return mat_index[id(m)]

        pos_chunks, uv_chunks, idx_chunks = [], [], []
        vbase_of = {}
        n_vertices = 0
        n_indices = 0
        any_uv = False
        node_rows = (A.NrbNodeDesc * max(1, len(nodes)))()
        for i, n in enumerate(nodes):
            r = node_rows[i]
            g = n.geometry
            r.shape = g.kind
            r.material = mat_id(n.material)
            r.param = (C.c_double * 3)(*g.param)
            r.rot = (C.c_double * 9)(*n.transform.rot.reshape(-1))
            r.trans = (C.c_double * 3)(*n.transform.trans)
            r.refr_coeff = n.refr_coeff
            r.refl_mix = n.refl_mix
            r.refl_atenuation = n.refl_atenuation
            r.alpha = n.alpha
            r.solid = 1 if n.solid else 0
            r.nmap_texture = tex_id(n.nmap)
            if g.kind == A.NRB_SHAPE_TRIMESH:
                key = g._coords_key
                if key not in vbase_of:
                    vbase_of[key] = n_vertices
                    pos_chunks.append(g.coords)
                    if g.uvs is not None:
                        any_uv = True
                        uv_chunks.append(g.uvs)
                    else:
                        uv_chunks.append(np.zeros((len(g.coords), 2), dtype=np.float32))  # src/obj.rs:383
                    n_vertices += len(g.coords)
                r.vertex_base = vbase_of[key]
                r.first_index = n_indices
                r.tri_count = len(g.faces)
                idx_chunks.append(g.faces.reshape(-1))
                n_indices += g.faces.size

        light_rows = (A.NrbLightDesc * max(1, len(lights)))()
        for i, l in enumerate(lights):
            light_rows[i].pos = (C.c_double * 3)(*l.pos)
            light_rows[i].radius = l.radius
            light_rows[i].racsample = l.racsample
            light_rows[i].color = (C.c_float * 3)(*l.color)

        self.node_rows = node_rows
        self.light_rows = light_rows
        self.mat_rows = (A.NrbMaterialDesc * max(1, len(mat_rows)))(*mat_rows)
        self.tex_rows = (A.NrbTextureDesc * max(1, len(tex_rows)))(*tex_rows)
        self.texels = (np.concatenate(texel_chunks) if texel_chunks else np.zeros((0, 4), np.float32)).astype(
            np.float32, copy=False)
        self.positions = (np.concatenate(pos_chunks) if pos_chunks else np.zeros((0, 3), np.float32))
        self.uvs = (np.concatenate(uv_chunks) if uv_chunks else np.zeros((0, 2), np.float32))
        self.indices = (np.concatenate(idx_chunks) if idx_chunks else np.zeros((0,), np.uint32)).astype(
            np.uint32, copy=False)
        self.positions = np.ascontiguousarray(self.positions, dtype=np.float32)
        self.uvs = np.ascontiguousarray(self.uvs, dtype=np.float32)
        self.texels = np.ascontiguousarray(self.texels, dtype=np.float32)
        self.n_triangles = n_indices // 3

        d = A.NrbSceneDesc()
        d.struct_size = C.sizeof(A.NrbSceneDesc)
        d.abi_version = A.NRB_ABI_VERSION
        d.n_nodes = len(nodes)
        d.n_lights = len(lights)
        d.n_materials = len(mat_rows)
        d.n_textures = len(tex_rows)
        d.nodes = C.cast(self.node_rows, C.POINTER(A.NrbNodeDesc))
        d.lights = C.cast(self.light_rows, C.POINTER(A.NrbLightDesc))
        d.materials = C.cast(self.mat_rows, C.POINTER(A.NrbMaterialDesc))
        d.textures = C.cast(self.tex_rows, C.POINTER(A.NrbTextureDesc))
        d.n_texels = len(self.texels)
        d.texels = self.texels.ctypes.data_as(C.POINTER(C.c_float))
        d.n_vertices = len(self.positions)
        d.positions = self.positions.ctypes.data_as(C.POINTER(C.c_float))
        d.uvs = self.uvs.ctypes.data_as(C.POINTER(C.c_float)) if (any_uv or len(self.uvs)) else None
        d.n_indices = len(self.indices)
        d.indices = self.indices.ctypes.data_as(C.POINTER(C.c_uint32))
        d.background = (C.c_float * 3)(*[float(x) for x in background])
        self.desc = d


def make_camera(width, height, ray_per_pixel, window_width, camera_eye, projection, seed=0, max_depth=0):
    """Pack the arguments of scene::render (src/scene.rs:29-36) into NrbCamera."""
    cam = A.NrbCamera()
    cam.width = int(width)
    cam.height = int(height)
    cam.ray_per_pixel = int(ray_per_pixel)
    cam.max_depth = int(max_depth)
    cam.window_width = float(window_width)
    cam.eye = (C.c_double * 3)(*[float(x) for x in camera_eye])
    P = np.asarray(projection, dtype=np.float64).reshape(4, 4)
    cam.projection = (C.c_double * 16)(*P.T.reshape(-1))  # column-major, nalgebra storage order
    cam.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return cam


class Image:
    """Image (src/image.rs:12-24): extents + row-major Vec<Vector3<f32>>."""

    def __init__(self, extents, pixels):
        self.extents = (int(extents[0]), int(extents[1]))
        self.pixels = np.asarray(pixels, dtype=np.float32).reshape(-1, 3)

    def as_array(self):
        w, h = self.extents
        return self.pixels.reshape(h, w, 3)

    def to_rgb8(self):
        """Image::to_png's quantisation (src/image.rs:64-77): clamp(c*255, 0, 255) truncated."""
        c = np.clip(self.pixels * np.float32(255.0), 0.0, 255.0)
        return c.astype(np.uint8).reshape(self.extents[1], self.extents[0], 3)

    def to_png(self, path):
        from PIL import Image as PILImage

        PILImage.fromarray(self.to_rgb8(), "RGB").save(path)


class Scene:
    """Scene (src/scene.rs:21-25).  Scene.new flattens and uploads (BVH build happens in the library)."""

    def __init__(self, nodes, lights, background=(1.0, 1.0, 1.0), device=0, upload=True, builder=None):
        self.nodes = list(nodes)
        self._lights = list(lights)
        self.background = tuple(float(x) for x in background)
        self.flat = FlatScene(self.nodes, self._lights, self.background)
        self.device = device
        self.builder = builder  # "sah" (host binned SAH), "lbvh" or "ploc" (device builds); None: library default / env NRB_BUILDER
        self._handle = None
        if upload:
            self.upload()

    new = None

    def upload(self):
        from . import _lib

        lib = _lib.load()
        h = C.c_void_p()
        if self.builder is None:
            _lib.check(lib.nrb_scene_create_opts(C.byref(self.flat.desc), int(self.device), None, C.byref(h)))
        else:
            opts = A.NrbBuildOptions({"lbvh": A.NRB_BUILDER_LBVH, "ploc": A.NRB_BUILDER_PLOC}.get(self.builder, A.NRB_BUILDER_SAH))
            _lib.check(lib.nrb_scene_create_opts(C.byref(self.flat.desc), int(self.device), C.byref(opts), C.byref(h)))
        self._handle = h

    def build_info(self):
        from . import _lib

        info = A.NrbBuildInfo()
        _lib.check(_lib.load().nrb_scene_build_info(self.handle, C.byref(info)))
        return info

    def lights(self):
        return self._lights

    def set_background(self, background):
        from . import _lib

        self.background = tuple(float(x) for x in background)
        self.flat.desc.background = (C.c_float * 3)(*self.background)
        if self._handle is not None:
            _lib.check(_lib.load().nrb_scene_set_background(self._handle, (C.c_float * 3)(*self.background)))

    @property
    def handle(self):
        if self._handle is None:
            raise RuntimeError("scene was not uploaded")
        return self._handle

    def close(self):
        if self._handle is not None:
            from . import _lib

            _lib.load().nrb_scene_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


Scene.new = staticmethod(lambda nodes, lights, background=(1.0, 1.0, 1.0), **k: Scene(nodes, lights, background, **k))


def render(scene, resolution, ray_per_pixel, window_width, camera_eye, projection, seed=0, max_depth=0,
           return_stats=False):
    """scene::render (src/scene.rs:29-116) through the C-ABI: host image out, synchronous."""
    from . import _lib

    lib = _lib.load()
    w, h = int(resolution[0]), int(resolution[1])
    cam = make_camera(w, h, ray_per_pixel, window_width, camera_eye, projection, seed, max_depth)
    out = np.empty((h * w, 3), dtype=np.float32)
    stats = A.NrbStats()
    _lib.check(lib.nrb_render(scene.handle, C.byref(cam), out.ctypes.data_as(C.POINTER(C.c_float)), C.byref(stats)))
    img = Image((w, h), out)
    return (img, stats) if return_stats else img
