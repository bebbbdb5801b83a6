"""Golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py from the f64 oracle)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_golden as G  # noqa: E402
import oracle_lib as O  # noqa: E402
from util import assert_parity, image_metrics  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def load(name):
    z = np.load(os.path.join(HERE, "golden", name + ".npz"))
    return z["image"], z["counts"]


@pytest.mark.parametrize("name", sorted(G.CASES))
def test_oracle_reproduces_golden(name):
    scene, _camd, cam, _ = G.build_case(name)
    img, st = O.OracleScene(scene.flat, 64).render(cam)
    gold, counts = load(name)
    np.testing.assert_allclose(img, gold, rtol=0, atol=2e-6)   # same binary: only libm / thread-order noise allowed
    assert [st.rays_primary, st.rays_reflect, st.rays_refract, st.rays_shadow, st.paths_truncated] == counts.tolist()


@pytest.mark.parametrize("name", sorted(G.CASES))
def test_f32_twin_close_to_golden(name):
    """The f32 mode (device twin) stays inside the path's stated tolerance of the f64 golden image."""
    scene, _camd, cam, _ = G.build_case(name)
    img, _ = O.OracleScene(scene.flat, 32).render(cam)
    gold, _ = load(name)
    assert_parity(img, gold, what="f32twin/" + name, wh=(cam.width, cam.height))


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(G.CASES))
def test_device_matches_golden(gpu, name):
    import ctypes as C

    from nrays_b200 import _abi as A
    from nrays_b200 import _lib

    scene, _camd, cam, (w, h, _spp, _win, _seed) = G.build_case(name, upload=True)
    out = np.empty((w * h, 3), np.float32)
    st = A.NrbStats()
    _lib.check(gpu.nrb_render(scene.handle, C.byref(cam), out.ctypes.data_as(C.POINTER(C.c_float)), C.byref(st)))
    gold, counts = load(name)
    m = assert_parity(out, gold, what="golden/" + name, wh=(w, h))
    assert m["mean_abs"] < 1e-3   # a few silhouette flips on a ~5 k-pixel image dominate the mean
    got = [st.rays_primary, st.rays_reflect, st.rays_refract, st.rays_shadow, st.paths_truncated]
    assert got[0] == counts[0]
    for a, b in zip(got, counts.tolist()):
        assert abs(a - b) <= max(4, 3e-3 * b), (name, got, counts.tolist())
    scene.close()
