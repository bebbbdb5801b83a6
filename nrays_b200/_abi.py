"""ctypes mirror of include/nrays_b200.h (the C-ABI boundary).

Field order and types must match the header exactly; tests/test_abi.py checks the struct
sizes against the values the shared library reports.
"""
import ctypes as C

NRB_ABI_VERSION = 1
NRB_TILE = 16

NRB_OK = 0
NRB_ERR_INVALID_ARG = 1
NRB_ERR_CUDA = 2
NRB_ERR_NO_DEVICE = 3
NRB_ERR_QUEUE_OVERFLOW = 4
NRB_ERR_UNSUPPORTED = 5
NRB_ERR_INTERNAL = 6

NRB_SHAPE_BALL = 0
NRB_SHAPE_CUBOID = 1
NRB_SHAPE_CYLINDER = 2
NRB_SHAPE_CAPSULE = 3
NRB_SHAPE_CONE = 4
NRB_SHAPE_PLANE = 5
NRB_SHAPE_TRIMESH = 6

NRB_MAT_PHONG = 0
NRB_MAT_NORMAL = 1
NRB_MAT_UV = 2

NRB_INTERP_BILINEAR = 0
NRB_INTERP_NEAREST = 1
NRB_OVERFLOW_CLAMP = 0
NRB_OVERFLOW_WRAP = 1


class NrbNodeDesc(C.Structure):
    _fields_ = [
        ("shape", C.c_int32),
        ("material", C.c_int32),
        ("param", C.c_double * 3),
        ("rot", C.c_double * 9),
        ("trans", C.c_double * 3),
        ("refr_coeff", C.c_double),
        ("refl_mix", C.c_float),
        ("refl_atenuation", C.c_float),
        ("alpha", C.c_float),
        ("solid", C.c_int32),
        ("nmap_texture", C.c_int32),
        ("_pad", C.c_int32),
        ("first_index", C.c_uint64),
        ("tri_count", C.c_uint64),
        ("vertex_base", C.c_uint64),
    ]


class NrbLightDesc(C.Structure):
    _fields_ = [
        ("pos", C.c_double * 3),
        ("radius", C.c_double),
        ("racsample", C.c_uint32),
        ("color", C.c_float * 3),
    ]


class NrbMaterialDesc(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("ambient", C.c_float * 3),
        ("diffuse", C.c_float * 3),
        ("specular", C.c_float * 3),
        ("shininess", C.c_float),
        ("texture", C.c_int32),
        ("alpha_texture", C.c_int32),
    ]


class NrbTextureDesc(C.Structure):
    _fields_ = [
        ("width", C.c_uint32),
        ("height", C.c_uint32),
        ("interpolation", C.c_int32),
        ("overflow", C.c_int32),
        ("texel_offset", C.c_uint64),
    ]


class NrbSceneDesc(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("abi_version", C.c_uint32),
        ("n_nodes", C.c_uint32),
        ("n_lights", C.c_uint32),
        ("n_materials", C.c_uint32),
        ("n_textures", C.c_uint32),
        ("nodes", C.POINTER(NrbNodeDesc)),
        ("lights", C.POINTER(NrbLightDesc)),
        ("materials", C.POINTER(NrbMaterialDesc)),
        ("textures", C.POINTER(NrbTextureDesc)),
        ("n_texels", C.c_uint64),
        ("texels", C.POINTER(C.c_float)),
        ("n_vertices", C.c_uint64),
        ("positions", C.POINTER(C.c_float)),
        ("uvs", C.POINTER(C.c_float)),
        ("n_indices", C.c_uint64),
        ("indices", C.POINTER(C.c_uint32)),
        ("background", C.c_float * 3),
        ("_pad", C.c_uint32),
    ]


class NrbCamera(C.Structure):
    _fields_ = [
        ("width", C.c_uint32),
        ("height", C.c_uint32),
        ("ray_per_pixel", C.c_uint32),
        ("max_depth", C.c_uint32),
        ("window_width", C.c_double),
        ("eye", C.c_double * 3),
        ("projection", C.c_double * 16),
        ("seed", C.c_uint64),
    ]


class NrbTileSet(C.Structure):
    _fields_ = [("first", C.c_uint32), ("stride", C.c_uint32)]


class NrbStats(C.Structure):
    _fields_ = [
        ("rays_primary", C.c_uint64),
        ("rays_reflect", C.c_uint64),
        ("rays_refract", C.c_uint64),
        ("rays_shadow", C.c_uint64),
        ("paths_truncated", C.c_uint64),
        ("waves", C.c_uint32),
        ("kernel_launches", C.c_uint32),
        ("ms_device", C.c_float),
        ("ms_trace", C.c_float),
        ("ms_shade", C.c_float),
        ("_pad", C.c_uint32),
        ("bvh_nodes", C.c_uint64),
        ("triangles", C.c_uint64),
        ("scene_bytes", C.c_uint64),
        ("launches_trace", C.c_uint32),
        ("launches_shade", C.c_uint32),
        ("rays_shadow_culled", C.c_uint64),
        ("ms_tail", C.c_float),
        ("ms_shade_kernel", C.c_float),
        ("launches_tail", C.c_uint32),
        ("_pad2", C.c_uint32),
        ("rays_tail", C.c_uint64),
    ]

    @property
    def rays_total(self):
        """BVH queries actually performed (the Mrays/s numerator)."""
        return self.rays_primary + self.rays_reflect + self.rays_refract + self.rays_shadow - self.rays_shadow_culled

    @property
    def rays_reference(self):
        """BVH queries of the reference semantics (it also casts the zero-weight light samples)."""
        return self.rays_primary + self.rays_reflect + self.rays_refract + self.rays_shadow

    def as_dict(self):
        d = {n: getattr(self, n) for n, _ in self._fields_ if not n.startswith("_")}
        d["rays_total"] = self.rays_total
        d["rays_reference"] = self.rays_reference
        return d


class NrbBuildInfo(C.Structure):
    _fields_ = [
        ("bvh_nodes", C.c_uint64),
        ("triangles", C.c_uint64),
        ("shapes", C.c_uint64),
        ("planes", C.c_uint64),
        ("transparent_candidates", C.c_uint64),
        ("max_depth", C.c_uint32),
        ("build_ms", C.c_float),
        ("gpu_build_ms", C.c_float),
        ("builder", C.c_uint32),
        ("node_format", C.c_uint32),
        ("_pad", C.c_uint32),
    ]


class NrbIpcHandle(C.Structure):
    _fields_ = [("bytes", C.c_ubyte * 64)]


class NrbBuildOptions(C.Structure):
    _fields_ = [("builder", C.c_uint32), ("_reserved", C.c_uint32 * 3)]


NRB_BUILDER_SAH = 0
NRB_BUILDER_LBVH = 1
NRB_BUILDER_PLOC = 2


# Every symbol include/nrays_b200.h declares (tests check the .so exports all of them).
EXPORTS = [
    "nrb_device_count",
    "nrb_scene_create",
    "nrb_scene_create_opts",
    "nrb_scene_build_info",
    "nrb_scene_validate",
    "nrb_scene_destroy",
    "nrb_scene_set_background",
    "nrb_scene_set_stream",
    "nrb_render",
    "nrb_render_device",
    "nrb_render_tiles_device",
    "nrb_render_tiles_to_image",
    "nrb_render_tiles_to_host",
    "nrb_render_tiles_to_image_rgb8",
    "nrb_render_tiles_to_host_rgb8",
    "nrb_ipc_alloc",
    "nrb_ipc_open",
    "nrb_ipc_close",
    "nrb_ipc_free",
    "nrb_host_register",
    "nrb_host_unregister",
    "nrb_tile_count",
    "nrb_tile_count_local",
    "nrb_untile_device",
    "nrb_render_rgb8",
    "nrb_host_alloc",
    "nrb_host_free",
    "nrb_last_error",
    "nrb_version",
]
