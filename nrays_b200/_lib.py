"""Loader for the CUDA shared library (nrays_b200/csrc/libnrays_b200.so).

There is no CPU fallback: if the library is missing or cannot be loaded the product path raises.
"""
import ctypes as C
import os

from . import _abi as A

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NRB_LIB") or os.path.join(_HERE, "csrc", "libnrays_b200.so")  # NRB_LIB: kernel-variant experiments
_lib = None


class NraysError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("nrays_b200 status %d: %s" % (status, message))
        self.status = status


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "CUDA library %s is missing — build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(make -C nrays_b200/csrc). There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp, u32, f32p = C.c_void_p, C.c_uint32, C.POINTER(C.c_float)
    lib.nrb_device_count.restype = C.c_int
    lib.nrb_scene_create.argtypes = [C.POINTER(A.NrbSceneDesc), C.c_int, C.POINTER(vp)]
    lib.nrb_scene_create.restype = C.c_int
    lib.nrb_scene_create_opts.argtypes = [C.POINTER(A.NrbSceneDesc), C.c_int, C.POINTER(A.NrbBuildOptions), C.POINTER(vp)]
    lib.nrb_scene_create_opts.restype = C.c_int
    lib.nrb_scene_build_info.argtypes = [vp, C.POINTER(A.NrbBuildInfo)]
    lib.nrb_scene_build_info.restype = C.c_int
    lib.nrb_scene_validate.argtypes = [C.POINTER(A.NrbSceneDesc), C.POINTER(A.NrbBuildInfo)]
    lib.nrb_scene_validate.restype = C.c_int
    lib.nrb_scene_destroy.argtypes = [vp]
    lib.nrb_scene_destroy.restype = None
    lib.nrb_scene_set_background.argtypes = [vp, C.POINTER(C.c_float)]
    lib.nrb_scene_set_background.restype = C.c_int
    lib.nrb_scene_set_stream.argtypes = [vp, vp]
    lib.nrb_scene_set_stream.restype = C.c_int
    lib.nrb_render.argtypes = [vp, C.POINTER(A.NrbCamera), f32p, C.POINTER(A.NrbStats)]
    lib.nrb_render.restype = C.c_int
    lib.nrb_render_device.argtypes = [vp, C.POINTER(A.NrbCamera), vp, C.POINTER(A.NrbStats)]
    lib.nrb_render_device.restype = C.c_int
    lib.nrb_render_tiles_device.argtypes = [vp, C.POINTER(A.NrbCamera), C.POINTER(A.NrbTileSet), vp,
                                            C.POINTER(u32), C.POINTER(A.NrbStats)]
    lib.nrb_render_tiles_device.restype = C.c_int
    lib.nrb_render_tiles_to_image.argtypes = [vp, C.POINTER(A.NrbCamera), C.POINTER(A.NrbTileSet), vp, C.POINTER(A.NrbStats)]
    lib.nrb_render_tiles_to_image.restype = C.c_int
    lib.nrb_render_tiles_to_host.argtypes = [vp, C.POINTER(A.NrbCamera), C.POINTER(A.NrbTileSet), vp, C.POINTER(A.NrbStats)]
    lib.nrb_render_tiles_to_host.restype = C.c_int
    lib.nrb_render_tiles_to_image_rgb8.argtypes = [vp, C.POINTER(A.NrbCamera), C.POINTER(A.NrbTileSet), vp, C.POINTER(A.NrbStats)]
    lib.nrb_render_tiles_to_image_rgb8.restype = C.c_int
    lib.nrb_render_tiles_to_host_rgb8.argtypes = [vp, C.POINTER(A.NrbCamera), C.POINTER(A.NrbTileSet), vp, C.POINTER(A.NrbStats)]
    lib.nrb_render_tiles_to_host_rgb8.restype = C.c_int
    lib.nrb_ipc_alloc.argtypes = [C.c_int, C.c_uint64, C.POINTER(vp), C.POINTER(A.NrbIpcHandle)]
    lib.nrb_ipc_alloc.restype = C.c_int
    lib.nrb_ipc_open.argtypes = [C.c_int, C.POINTER(A.NrbIpcHandle), C.POINTER(vp)]
    lib.nrb_ipc_open.restype = C.c_int
    lib.nrb_ipc_close.argtypes = [C.c_int, vp]
    lib.nrb_ipc_close.restype = C.c_int
    lib.nrb_ipc_free.argtypes = [C.c_int, vp]
    lib.nrb_ipc_free.restype = C.c_int
    lib.nrb_host_register.argtypes = [C.c_int, vp, C.c_uint64, C.POINTER(vp)]
    lib.nrb_host_register.restype = C.c_int
    lib.nrb_host_unregister.argtypes = [C.c_int, vp]
    lib.nrb_host_unregister.restype = C.c_int
    lib.nrb_tile_count.argtypes = [u32, u32]
    lib.nrb_tile_count.restype = u32
    lib.nrb_tile_count_local.argtypes = [u32, u32, C.POINTER(A.NrbTileSet)]
    lib.nrb_tile_count_local.restype = u32
    lib.nrb_untile_device.argtypes = [C.c_int, vp, vp, u32, u32, u32, u32, vp]
    lib.nrb_untile_device.restype = C.c_int
    lib.nrb_render_rgb8.argtypes = [vp, C.POINTER(A.NrbCamera), C.POINTER(C.c_uint8), C.POINTER(A.NrbStats)]
    lib.nrb_render_rgb8.restype = C.c_int
    lib.nrb_host_alloc.argtypes = [C.c_uint64]
    lib.nrb_host_alloc.restype = vp
    lib.nrb_host_free.argtypes = [vp]
    lib.nrb_host_free.restype = None
    lib.nrb_last_error.restype = C.c_char_p
    lib.nrb_version.restype = C.c_char_p
    _lib = lib
    return lib


def check(status):
    if status != A.NRB_OK:
        raise NraysError(status, load().nrb_last_error().decode("utf-8", "replace"))
