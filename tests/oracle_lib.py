"""ctypes binding of the CPU oracle (oracle/libnrays_oracle.so).  TEST INFRASTRUCTURE: only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this."""
import ctypes as C
import os
import subprocess

import numpy as np

from nrays_b200 import _abi as A

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(_ROOT, "oracle", "libnrays_oracle.so")
_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        subprocess.check_call(["make", "-C", os.path.join(_ROOT, "oracle")])
    lib = C.CDLL(LIB_PATH)
    vp, dp = C.c_void_p, C.POINTER(C.c_double)
    lib.nro_scene_create.argtypes = [C.POINTER(A.NrbSceneDesc), C.c_int, C.POINTER(vp)]
    lib.nro_scene_destroy.argtypes = [vp]
    lib.nro_scene_destroy.restype = None
    lib.nro_render.argtypes = [vp, C.POINTER(A.NrbCamera), C.c_int, C.c_uint64, C.c_uint64, C.POINTER(C.c_float),
                               C.POINTER(A.NrbStats)]
    lib.nro_cast.argtypes = [vp, C.c_uint32, dp, dp, dp, dp, dp, C.POINTER(C.c_int)]
    lib.nro_trace.argtypes = [vp, C.POINTER(A.NrbCamera), dp, dp, C.c_uint32, C.c_uint32, C.POINTER(C.c_float)]
    lib.nro_intersects_ray.argtypes = [vp, dp, dp, C.c_double, C.POINTER(C.c_float)]
    lib.nro_texture_sample.argtypes = [vp, C.c_uint32, C.c_double, C.c_double, C.POINTER(C.c_float)]
    lib.nro_aabb_toi.argtypes = [dp, dp, dp, dp, C.c_int, dp]
    lib.nro_primary_ray.argtypes = [C.POINTER(A.NrbCamera), C.c_uint32, C.c_uint32, dp, dp]
    lib.nro_philox4x32_10.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    lib.nro_philox4x32_10.restype = None
    lib.nro_last_error.restype = C.c_char_p
    _lib = lib
    return lib


def _d3(v):
    return (C.c_double * 3)(*[float(x) for x in v])


class OracleScene:
    """Oracle twin of Scene: built from the same flattened tables (FlatScene.desc)."""

    def __init__(self, flat, bits=64):
        self.flat = flat  # keeps the buffers alive
        self.bits = bits
        self.h = C.c_void_p()
        st = load().nro_scene_create(C.byref(flat.desc), bits, C.byref(self.h))
        if st != 0:
            raise RuntimeError("oracle: %s" % load().nro_last_error().decode())

    def close(self):
        if self.h:
            load().nro_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def render(self, cam, threads=0, first_pixel=0, n_pixels=0, out=None):
        n = cam.width * cam.height
        if out is None:
            out = np.zeros((n, 3), dtype=np.float32)
        stats = A.NrbStats()
        st = load().nro_render(self.h, C.byref(cam), threads, first_pixel, n_pixels,
                               out.ctypes.data_as(C.POINTER(C.c_float)), C.byref(stats))
        if st != 0:
            raise RuntimeError("oracle status %d: %s" % (st, load().nro_last_error().decode()))
        return out, stats

    def cast(self, node, o, d):
        toi = C.c_double()
        n = (C.c_double * 3)()
        uv = (C.c_double * 2)()
        has = C.c_int()
        hit = load().nro_cast(self.h, node, _d3(o), _d3(d), C.byref(toi), n, uv, C.byref(has))
        if not hit:
            return None
        return dict(toi=toi.value, normal=np.array(n[:]), uv=(np.array(uv[:]) if has.value else None))

    def trace(self, cam, o, d, pixel=0, sample=0):
        rgb = (C.c_float * 3)()
        load().nro_trace(self.h, C.byref(cam), _d3(o), _d3(d), pixel, sample, rgb)
        return np.array(rgb[:], dtype=np.float32)

    def intersects_ray(self, o, d, maxtoi):
        f = (C.c_float * 3)()
        some = load().nro_intersects_ray(self.h, _d3(o), _d3(d), float(maxtoi), f)
        return np.array(f[:], dtype=np.float32) if some else None

    def texture_sample(self, tex, u, v):
        rgba = (C.c_float * 4)()
        st = load().nro_texture_sample(self.h, tex, float(u), float(v), rgba)
        if st != 0:
            raise RuntimeError(load().nro_last_error().decode())
        return np.array(rgba[:], dtype=np.float32)


def aabb_toi(mins, maxs, o, d, solid=True):
    t = C.c_double()
    hit = load().nro_aabb_toi(_d3(mins), _d3(maxs), _d3(o), _d3(d), 1 if solid else 0, C.byref(t))
    return t.value if hit else None


def primary_ray(cam, pixel, sample=0):
    o = (C.c_double * 3)()
    d = (C.c_double * 3)()
    st = load().nro_primary_ray(C.byref(cam), pixel, sample, o, d)
    if st != 0:
        return None
    return np.array(o[:]), np.array(d[:])


def philox(ctr, key):
    out = (C.c_uint32 * 4)()
    load().nro_philox4x32_10((C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), out)
    return [int(x) for x in out]
