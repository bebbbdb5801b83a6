"""GPU parity tests proper: the CUDA path (through the C-ABI) against the CPU oracle on the same
flattened scene, seeded inputs, sizes the oracle finishes in seconds.  Tolerance: per-channel
|delta| <= 1/255 on >= 99.9 % of pixels (util.TOL / MAX_FRAC_OVER; f32 device vs f64 reference)."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as O
from nrays_b200 import (Ball, Capsule, Cone, Cuboid, Cylinder, ImageData, Interpolation, Isometry3, Light,
                        NormalMaterial, Overflow, PhongMaterial, Plane, Scene, SceneNode, Texture2d, TriMesh, UVMaterial,
                        _abi as A, _lib, camera_projection, configs, make_camera, render)
from util import (TOL, assert_counts_close, assert_parity, checker_texture, default_phong, image_metrics, look, node, quad_mesh,
                  render_both)

pytestmark = pytest.mark.gpu


class _Env:
    def __init__(self, **kv):
        self.kv = {k: str(v) for k, v in kv.items()}

    def __enter__(self):
        import os
        self.old = {k: os.environ.get(k) for k in self.kv}
        os.environ.update(self.kv)

    def __exit__(self, *a):
        import os
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def phong(ka=(0.1, 0.1, 0.1), kd=(0.8, 0.7, 0.6), ks=(0.5, 0.5, 0.5), ns=40.0, tex=None, amap=None):
    return PhongMaterial(ka, kd, ks, tex, amap, ns)


def zoo(materials):
    shapes = [Ball(0.8), Cuboid((0.6, 0.5, 0.7)), Cylinder(0.7, 0.5), Capsule(0.5, 0.35), Cone(0.8, 0.6)]
    nodes = []
    for i, (g, m) in enumerate(zip(shapes, materials)):
        nodes.append(node(g, m, pos=(-3.2 + 1.6 * i, 0.2 * (i % 2), 0.3 * i), angle=(20 * i, 35 * i, 10 * i)))
    nodes.append(node(Plane((0, 1, 0)), phong(), pos=(0, -1.2, 0)))
    return nodes


def test_shape_zoo_phong(gpu):
    tex = Texture2d(checker_texture(16, 8), Interpolation.Bilinear, Overflow.Wrap)
    mats = [phong(tex=tex), phong(tex=tex), phong(), phong(kd=(0.2, 0.9, 0.3)), phong(ks=(0, 0, 0))]
    lights = [Light((2, 4, -3), 0.0, 1, (1.0, 0.9, 0.8)), Light((-3, 2, -2), 0.0, 1, (0.3, 0.3, 0.5))]
    img, st, ref, ost = render_both(zoo(mats), lights, eye=(0.5, 2.0, -7.0), w=160, h=96)
    assert_parity(img, ref, what="zoo/phong", wh=(160, 96))
    assert_counts_close(st, ost)


def test_shape_zoo_debug_materials_and_texture_modes(gpu):
    t1 = Texture2d(checker_texture(5, 9, 4), Interpolation.Nearest, Overflow.ClampToEdges)
    t2 = Texture2d(checker_texture(8, 8, 5), Interpolation.Bilinear, Overflow.ClampToEdges)
    mats = [UVMaterial(), NormalMaterial(), UVMaterial(), NormalMaterial(), phong(tex=t1)]
    nodes = zoo(mats) + [node(Cuboid((0.4, 0.4, 0.4)), phong(tex=t2), pos=(0, 1.5, 1.0), angle=(45, 0, 30))]
    img, st, ref, ost = render_both(nodes, [Light((0, 5, -4), 0.0, 1, (1, 1, 1))], eye=(0.0, 1.5, -7.5), w=160, h=96)
    assert_parity(img, ref, what="zoo/debug", wh=(160, 96))
    assert_counts_close(st, ost)


def test_transparency_refraction_reflection_chain(gpu):
    """alpha < 1 on every shape kind + reflective floor: refraction enters/exits, transparent shadows."""
    glass = [phong(ka=(0.2, 0.1, 0.1)), phong(ka=(0.1, 0.2, 0.1)), phong(ka=(0.1, 0.1, 0.2)), phong(), phong()]
    nodes = []
    shapes = [Ball(0.8), Cuboid((0.6, 0.5, 0.7)), Cylinder(0.7, 0.5), Capsule(0.5, 0.35), Cone(0.8, 0.6)]
    for i, (g, m) in enumerate(zip(shapes, glass)):
        nodes.append(node(g, m, pos=(-3.2 + 1.6 * i, 0.0, 0.0), angle=(0, 15 * i, 0), alpha=0.3 + 0.1 * i, refr=1.3))
    nodes.append(node(Plane((0, 1, 0)), phong(), pos=(0, -1.2, 0), refl=(0.3, 0.4)))
    img, st, ref, ost = render_both(nodes, [Light((0, 5, -2), 0.0, 1, (1, 1, 1))], eye=(0.0, 2.0, -7.0), w=160, h=96)
    assert_parity(img, ref, what="glass", wh=(160, 96))
    assert_counts_close(st, ost)
    assert st.rays_refract > 0 and st.rays_reflect > 0


def test_area_light_and_jitter_share_the_rng(gpu):
    """Philox-keyed jitter + cube light samples (src/light.rs:56-63): same seed -> same image on both sides."""
    nodes = [node(Ball(1.0), phong(), pos=(0, 0, 0)), node(Plane((0, 1, 0)), phong(), pos=(0, -1, 0))]
    lights = [Light((1.5, 3, -1), 0.6, 10, (1, 1, 1))]  # 9 samples
    img, st, ref, ost = render_both(nodes, lights, eye=(0, 1.5, -5), w=96, h=64, spp=3, window=1.0, seed=11)
    assert_parity(img, ref, what="area light", wh=(96, 64))
    assert st.rays_shadow == ost.rays_shadow or abs(int(st.rays_shadow) - int(ost.rays_shadow)) < 0.002 * ost.rays_shadow
    img2, _, _, _ = render_both(nodes, lights, eye=(0, 1.5, -5), w=96, h=64, spp=3, window=1.0, seed=12)
    assert np.abs(img2 - img).max() > 1e-3  # a different seed gives a different frame


def test_alpha_mapped_mesh_and_per_node_shadow_semantics(gpu):
    """Opacity-mapped quads over a floor: shadows are filtered per SceneNode closest hit (F10, A.6)."""
    rng = np.random.default_rng(0)
    px = np.ones((16 * 16, 4), np.float32)
    px[:, 3] = (rng.uniform(size=256) > 0.5).astype(np.float32)
    amap = Texture2d(ImageData(px, (16, 16)), Interpolation.Nearest, Overflow.Wrap)
    P, F, UV = quad_mesh(1.5, 6, y=1.0)
    P2 = P.copy()
    P2[:, 1] = 0.5
    two_layers = TriMesh(np.concatenate([P, P2]), np.concatenate([F, F + len(P)]), np.concatenate([UV, UV * 2]))
    Pf, Ff, UVf = quad_mesh(4.0, 8, y=-0.5)
    nodes = [node(two_layers, phong(ka=(0.3, 0.3, 0.3), amap=amap)), node(TriMesh(Pf, Ff, UVf), phong()),
             node(TriMesh(P + np.float32([0, 1.2, 0]), F, UV), phong(ka=(0.5, 0.2, 0.2)), alpha=0.4)]
    # eye slightly off the symmetry axes: from (0, 3, -6) the centre row of the image runs exactly along the mesh's shared
    # edge z = -1 (an exact tie between two triangles on a Nearest-sampled texel boundary: order-dependent in the reference
    # itself; exact ties have their own test below)
    img, st, ref, ost = render_both(nodes, [Light((0.5, 6, -0.5), 0.0, 1, (1, 1, 1))], eye=(0.07, 3.03, -6.0), w=160, h=120)
    assert_parity(img, ref, what="alpha map", wh=(160, 120), twin=render_both.twin)
    assert_counts_close(st, ost)   # rays_shadow counts the reference's light samples, cast or not
    # hits on fully transparent texels carry weight 0: their light samples add exactly nothing and are not cast
    assert 0 < st.rays_shadow_culled < st.rays_shadow and ost.rays_shadow_culled == 0
    assert st.rays_total == st.rays_reference - st.rays_shadow_culled


def test_exact_ties_resolve_to_one_of_the_tied_surfaces(gpu):
    """Two coplanar quads in different SceneNodes (exact tie in toi) and a ray grid that also runs along shared triangle edges:
    best_first_search keeps the first found (strict <), so the winner is order-dependent in the reference itself —
    the device must show ONE of the tied surfaces' colours per pixel, never a mixture, never the background."""
    quad = np.array([[-1, -1, 0], [1, -1, 0], [1, 1, 0], [-1, 1, 0]], np.float32)
    Fc = np.array([[0, 1, 2], [0, 2, 3]], np.uint32)
    red = PhongMaterial((1, 0, 0), (0, 0, 0), (0, 0, 0), None, None, 1.0)
    green = PhongMaterial((0, 1, 0), (0, 0, 0), (0, 0, 0), None, None, 1.0)
    nodes = [node(TriMesh(quad, Fc, None), red), node(TriMesh(quad.copy(), Fc, None), green)]
    scene = Scene(nodes, [], (0.0, 0.0, 1.0))
    w = h = 64
    eye = (0.0, 0.0, -3.0)
    img = render(scene, (w, h), 1, 0.0, eye, camera_projection(eye, (0, 0, 0), 30.0, w, h)).pixels   # quad fills the frame
    scene.close()
    is_red = np.all(np.abs(img - np.float32([1, 0, 0])) < 1e-6, axis=1)
    is_green = np.all(np.abs(img - np.float32([0, 1, 0])) < 1e-6, axis=1)
    assert (is_red | is_green).all(), "a tied pixel is neither of the tied surfaces"


def _shift_texture(lo, hi, seed, size=(8, 8)):
    rng = np.random.default_rng(seed)
    px = np.ones((size[0] * size[1], 4), np.float32)
    px[:, :3] = rng.uniform(lo, hi, (size[0] * size[1], 3)).astype(np.float32)
    return Texture2d(ImageData(px, size), Interpolation.Bilinear, Overflow.Wrap)


def test_nmap_depth_shift_nodes(gpu):
    """SceneNode.nmap (src/scene_node.rs:60-70): toi -= mean(nmap.sample(uv).rgb) after the node's own cast, on a ball, a
    rotated cuboid and a mesh, next to plain nodes, with shadows, a reflective plane and a transparent shifted node —
    the shifted toi decides closest hits, hit points and shadow occlusion, in the reference's best-first order."""
    from nrays_b200 import Isometry3, SceneNode

    def shifted(geom, mat, nmap, pos, angle=(0, 0, 0), alpha=1.0):
        return SceneNode(mat, 0.0, 0.0, alpha, 1.0, Isometry3.new(pos, np.radians(angle)), geom, nmap, False)

    P, F, UV = quad_mesh(1.2, 5, y=0.0)
    nodes = [shifted(Ball(0.9), phong(ka=(0.2, 0.1, 0.1)), _shift_texture(0.0, 0.6, 1), (-2.2, 0.2, 0.0)),
             shifted(Cuboid((0.7, 0.6, 0.5)), phong(ka=(0.1, 0.2, 0.1)), _shift_texture(0.1, 0.4, 2), (0.0, 0.1, 0.4), (20, 35, 10)),
             shifted(TriMesh(P, F, UV), phong(ka=(0.1, 0.1, 0.3)), _shift_texture(0.0, 0.8, 3, (5, 7)), (2.3, 0.3, 0.2), (-50, 0, 15)),
             shifted(Ball(0.5), phong(ka=(0.3, 0.3, 0.1)), _shift_texture(0.2, 0.3, 4), (0.9, 1.4, -0.8), alpha=0.5),
             shifted(Cylinder(0.6, 0.4), phong(), _shift_texture(0.5, 0.9, 5), (-0.9, 1.3, 0.9)),   # no uvs: the texture is inert
             node(Ball(0.6), phong(), pos=(-0.2, -0.2, -1.6)), node(Cuboid((0.4, 0.4, 0.4)), NormalMaterial(), pos=(1.6, -0.4, -1.2)),
             node(Plane((0, 1, 0)), phong(), pos=(0, -1.2, 0), refl=(0.3, 0.5))]
    lights = [Light((1.5, 5.0, -3.0), 0.0, 1, (1, 1, 1)), Light((-3.0, 4.0, -2.0), 0.3, 4, (0.6, 0.6, 0.9))]
    img, st, ref, ost = render_both(nodes, lights, eye=(0.3, 1.6, -6.5), w=160, h=112, spp=2, window=1.0, seed=3)
    assert_parity(img, ref, what="nmap", wh=(160, 112), twin=render_both.twin)
    assert_counts_close(st, ost)
    # and the shift is really in the picture: the same scene without the textures differs
    plain = [SceneNode(n.material, n.refl_mix, n.refl_atenuation, n.alpha, n.refr_coeff, n.transform, n.geometry, None, n.solid) for n in nodes]
    img0, _, _, _ = render_both(plain, lights, eye=(0.3, 1.6, -6.5), w=160, h=112, spp=2, window=1.0, seed=3)
    assert (np.abs(img0 - img).max(axis=1) > 0.02).mean() > 0.02
    # all driver paths agree on it (tail / no tail / tiny batches)
    for env in (dict(NRB_TAIL_RAYS=0), dict(NRB_TAIL_RAYS=1 << 30), dict(NRB_BATCH_SLOTS=4096, NRB_SHADOW_CAP=2048)):
        with _Env(**env):
            img2, st2, _, _ = render_both(nodes, lights, eye=(0.3, 1.6, -6.5), w=160, h=112, spp=2, window=1.0, seed=3)
        np.testing.assert_allclose(img2, img, rtol=0, atol=3e-5, err_msg=str(env))
        assert st2.rays_reference == st.rays_reference


def test_solid_flag_and_camera_inside_objects(gpu):
    nodes = [node(Ball(3.0), phong(), pos=(0, 0, 0), solid=False), node(Cuboid((0.5, 0.5, 0.5)), NormalMaterial(), pos=(0, 0, 1.5)),
             node(Cylinder(0.4, 0.3), UVMaterial(), pos=(1.2, 0, 1.0))]
    img, st, ref, ost = render_both(nodes, [Light((0, 1, 0), 0.0, 1, (1, 1, 1))], eye=(0, 0, -1.0), at=(0, 0, 1), w=96, h=64)
    assert_parity(img, ref, what="inside ball", wh=(96, 64))
    nodes[0] = node(Ball(3.0), phong(), pos=(0, 0, 0), solid=True)   # toi 0 everywhere
    img, st, ref, ost = render_both(nodes, [Light((0, 1, 0), 0.0, 1, (1, 1, 1))], eye=(0, 0, -1.0), at=(0, 0, 1), w=96, h=64)
    assert_parity(img, ref, what="solid ball", wh=(96, 64))


def test_transparent_plane_candidate_and_plane_only_scene(gpu):
    nodes = [node(Plane((0, 0, -1)), phong(ka=(0.2, 0.3, 0.4)), pos=(0, 0, 2), alpha=0.5),
             node(Plane((0, 1, 0)), phong(), pos=(0, -1, 0)), node(Ball(0.7), phong(), pos=(0.3, 0, 4))]
    img, st, ref, ost = render_both(nodes, [Light((0, 3, -3), 0.0, 1, (1, 1, 1))], eye=(0, 1, -4), w=96, h=64)
    assert_parity(img, ref, what="transparent plane", wh=(96, 64))
    assert_counts_close(st, ost)
    img, st, ref, ost = render_both(nodes[:2], [Light((0, 3, -3), 0.0, 1, (1, 1, 1))], eye=(0, 1, -4), w=64, h=48)
    assert_parity(img, ref, what="planes only", wh=(64, 48))


def test_depth_cap_matches_oracle(gpu):
    mir = [node(Plane((0, 1, 0)), NormalMaterial(), pos=(0, -1, 0), refl=(0.6, 0.0)),
           node(Plane((0, -1, 0)), NormalMaterial(), pos=(0, 1, 0), refl=(0.6, 0.0)),
           node(Plane((0, 0, -1)), UVMaterial(), pos=(0, 0, 6))]
    img, st, ref, ost = render_both(mir, [], eye=(0, 0.2, -3), at=(0, 0.1, 0), w=64, h=48, max_depth=9)
    assert_parity(img, ref, what="mirror box", wh=(64, 48))
    assert st.paths_truncated == ost.paths_truncated > 0
    assert st.rays_reflect == ost.rays_reflect


def test_empty_scene_and_background(gpu):
    img, st, ref, ost = render_both([], [], eye=(0, 0, -3), w=33, h=17, background=(0.25, 0.5, 0.75))
    np.testing.assert_allclose(img, np.tile(np.float32([0.25, 0.5, 0.75]), (33 * 17, 1)))
    np.testing.assert_allclose(ref, img)
    assert st.rays_primary == 33 * 17 and st.rays_total == 33 * 17
    scene = Scene([node(Ball(1.0), NormalMaterial(), pos=(0, 0, 50))], [], (1, 1, 1))
    scene.set_background((0.1, 0.2, 0.3))
    out = render(scene, (8, 8), 1, 0.0, (0, 0, 0), camera_projection((0, 0, 0), (0, 1, 0.01), 30, 8, 8))
    np.testing.assert_allclose(out.pixels, np.tile(np.float32([0.1, 0.2, 0.3]), (64, 1)))
    scene.close()


@pytest.mark.parametrize("w,h,spp", [(1, 1, 1), (17, 9, 2), (16, 16, 1), (31, 33, 5)])
def test_ragged_resolutions(gpu, w, h, spp):
    nodes = [node(Ball(1.0), phong(), pos=(0, 0, 0)), node(Plane((0, 1, 0)), NormalMaterial(), pos=(0, -1, 0))]
    img, st, ref, ost = render_both(nodes, [Light((2, 3, -2), 0.0, 1, (1, 1, 1))], eye=(0, 1, -4), w=w, h=h, spp=spp, window=1.0, seed=4)
    assert st.rays_primary == w * h * spp
    assert_parity(img, ref, what="ragged %dx%dx%d" % (w, h, spp), wh=(w, h))


def test_error_codes(gpu):
    lib = gpu
    scene = Scene([node(Ball(1.0), NormalMaterial())], [], (1, 1, 1))
    out = np.zeros((16, 3), np.float32)
    outp = out.ctypes.data_as(C.POINTER(C.c_float))
    P = camera_projection((0, 0, -3), (0, 0, 0), 45, 4, 4)
    cam = make_camera(4, 4, 0, 0.0, (0, 0, -3), P)               # assert!(ray_per_pixel > 0): src/scene.rs:37
    assert lib.nrb_render(scene.handle, C.byref(cam), outp, None) == A.NRB_ERR_INVALID_ARG
    assert b"ray_per_pixel" in lib.nrb_last_error()
    cam = make_camera(0, 4, 1, 0.0, (0, 0, -3), P)
    assert lib.nrb_render(scene.handle, C.byref(cam), outp, None) == A.NRB_ERR_INVALID_ARG
    cam = make_camera(4, 4, 1, 0.0, (0, 0, -3), np.zeros((4, 4)))  # w == 0: from_homogeneous().unwrap() panics
    assert lib.nrb_render(scene.handle, C.byref(cam), outp, None) == A.NRB_ERR_INVALID_ARG
    assert lib.nrb_render(scene.handle, None, outp, None) == A.NRB_ERR_INVALID_ARG
    scene.close()
    # malformed tables
    flat = Scene([node(Ball(1.0), NormalMaterial())], [], upload=False).flat
    flat.node_rows[0].material = 7
    h = C.c_void_p()
    assert lib.nrb_scene_create(C.byref(flat.desc), 0, C.byref(h)) == A.NRB_ERR_INVALID_ARG
    flat = Scene([node(Ball(1.0), NormalMaterial())], [], upload=False).flat
    flat.node_rows[0].nmap_texture = 3      # no such texture
    assert lib.nrb_scene_create(C.byref(flat.desc), 0, C.byref(h)) == A.NRB_ERR_INVALID_ARG
    Pm, Fm, UVm = quad_mesh(1.0, 1)
    flat = Scene([node(TriMesh(Pm, Fm, UVm), NormalMaterial())], [], upload=False).flat
    flat.indices[0] = 1000
    assert lib.nrb_scene_create(C.byref(flat.desc), 0, C.byref(h)) == A.NRB_ERR_INVALID_ARG
    flat = Scene([node(Ball(1.0), NormalMaterial())], [], upload=False).flat
    assert lib.nrb_scene_create(C.byref(flat.desc), 99, C.byref(h)) == A.NRB_ERR_INVALID_ARG
    flat.desc.abi_version = 99
    assert lib.nrb_scene_create(C.byref(flat.desc), 0, C.byref(h)) == A.NRB_ERR_INVALID_ARG


def test_rgb8_output_is_the_png_quantisation(gpu):
    scene, camd, cfg = configs.build("C1", globe_size=(32, 16))
    w = h = 64
    cam = make_camera(w, h, 1, 0.0, camd.eye, camd.projection((w, h)))
    f = np.empty((w * h, 3), np.float32)
    b = np.empty((w * h, 3), np.uint8)
    _lib.check(gpu.nrb_render(scene.handle, C.byref(cam), f.ctypes.data_as(C.POINTER(C.c_float)), None))
    _lib.check(gpu.nrb_render_rgb8(scene.handle, C.byref(cam), b.ctypes.data_as(C.POINTER(C.c_uint8)), None))
    exp = np.clip(f * np.float32(255.0), 0, 255).astype(np.uint8)       # src/image.rs:64-77
    assert (np.abs(exp.astype(int) - b.astype(int)) > 1).mean() == 0     # atomics order can flip a +-1 boundary
    assert (exp != b).mean() < 0.01
    scene.close()


def test_tiles_resolved_into_registered_host_memory(gpu):
    """nrb_host_register maps caller-owned host memory (in production: a shared-memory segment all ranks map); the
    resolve kernels of all (virtual) ranks store their tiles straight into that host image."""
    import mmap

    from nrays_b200 import dist

    scene, camd, cfg = configs.build("C3", target_tris=30000, lod=4)
    w, h, world = 192, 100, 4       # 12 tile columns: 3 per rank; the last tile row is ragged (100 = 6 * 16 + 4)
    cam = make_camera(w, h, 2, 1.0, camd.eye, camd.projection((w, h)), seed=9)
    full = np.empty(w * h * 3, np.float32)
    _lib.check(gpu.nrb_render(scene.handle, C.byref(cam), full.ctypes.data_as(C.POINTER(C.c_float)), None))
    buf = mmap.mmap(-1, w * h * 12)          # page-aligned anonymous mapping, like a shared-memory segment
    host = np.frombuffer(buf, dtype=np.float32)
    host[:] = -5.0
    cbuf = C.c_char.from_buffer(buf)
    dptr = C.c_void_p()
    _lib.check(gpu.nrb_host_register(0, C.c_void_p(C.addressof(cbuf)), w * h * 12, C.byref(dptr)))
    for r in range(world):
        dist.render_tiles_to_image(scene, cam, r, world, dptr)
    np.testing.assert_allclose(host, full, rtol=0, atol=3e-5)
    # the DMA form: each (virtual) rank drops its tile columns into the host image with one strided 2-D copy
    host[:] = -6.0
    for r in range(world):
        dist.render_tiles_to_host(scene, cam, r, world, C.addressof(cbuf))
    np.testing.assert_allclose(host, full, rtol=0, atol=3e-5)
    ts = A.NrbTileSet(0, 5)        # 12 tile columns do not split over 5 ranks: the caller must use the other exchange
    assert gpu.nrb_render_tiles_to_host(scene.handle, C.byref(cam), C.byref(ts), C.c_void_p(C.addressof(cbuf)), None) == A.NRB_ERR_UNSUPPORTED
    _lib.check(gpu.nrb_host_unregister(0, C.c_void_p(C.addressof(cbuf))))
    assert gpu.nrb_host_register(0, None, 16, C.byref(dptr)) == A.NRB_ERR_INVALID_ARG
    del host, cbuf
    buf.close()
    scene.close()


def test_pinned_destination_overlaps_the_copy_and_patches_the_tail(gpu):
    """nrb_render into PINNED host memory starts the image's device->host copy when the frame enters its tail phase and
    then stores the pixels the tail changed straight into the (mapped) host image; the result is the image of the plain path."""
    scene, camd, cfg = configs.build("C3", target_tris=30000, lod=4)   # alpha-mapped foliage: a real tail phase
    w, h = 208, 117
    cam = make_camera(w, h, 3, 1.0, camd.eye, camd.projection((w, h)), seed=11)
    n = w * h * 3

    def go(ptr):
        st = A.NrbStats()
        _lib.check(gpu.nrb_render(scene.handle, C.byref(cam), C.cast(ptr, C.POINTER(C.c_float)), C.byref(st)))
        return st

    pageable = np.full(n, -1.0, np.float32)
    st_a = go(pageable.ctypes.data)
    hp = gpu.nrb_host_alloc(n * 4)
    pinned = np.ctypeslib.as_array(C.cast(hp, C.POINTER(C.c_float)), shape=(n,))
    pinned[:] = -2.0
    st_b = go(hp)
    np.testing.assert_allclose(pinned, pageable, rtol=0, atol=3e-5)
    assert st_b.rays_total == st_a.rays_total
    assert st_b.kernel_launches == st_a.kernel_launches + 1      # early resolve + patch kernel instead of one resolve
    with _Env(NRB_EARLY_COPY=0):
        pinned[:] = -3.0
        st_c = go(hp)
    np.testing.assert_allclose(pinned, pageable, rtol=0, atol=3e-5)
    assert st_c.kernel_launches == st_a.kernel_launches
    # a frame without a tail phase (one opaque ball, no lights) takes the plain path even with a pinned destination
    scene.close()
    plain = Scene([node(Ball(1.0), NormalMaterial())], [], (0.2, 0.3, 0.4))
    cam2 = make_camera(64, 48, 1, 0.0, (0, 0, -5), look((0, 0, -5), (0, 0, 0), 45.0, 64, 48), seed=0)
    st = A.NrbStats()
    _lib.check(gpu.nrb_render(plain.handle, C.byref(cam2), C.cast(hp, C.POINTER(C.c_float)), C.byref(st)))
    ref = np.empty(64 * 48 * 3, np.float32)
    _lib.check(gpu.nrb_render(plain.handle, C.byref(cam2), ref.ctypes.data_as(C.POINTER(C.c_float)), None))
    np.testing.assert_array_equal(pinned[:64 * 48 * 3], ref)
    plain.close()
    gpu.nrb_host_free(hp)


def test_tile_sharded_render_equals_full_frame(gpu):
    """8 virtual ranks on one GPU: packed tiles + un-tile == the unsharded frame (RNG keyed by global pixel)."""
    import torch

    from nrays_b200 import dist

    scene, camd, cfg = configs.build("C3", target_tris=30000, lod=4)
    w, h, world = 200, 120, 8
    cam = make_camera(w, h, 2, 1.0, camd.eye, camd.projection((w, h)), seed=5)
    full = torch.empty(h * w * 3, dtype=torch.float32, device="cuda")
    st_full = dist.render_device(scene, cam, full)
    tpr = dist.tiles_per_rank(w, h, world)
    packed = torch.zeros((world, tpr, 16, 16, 3), dtype=torch.float32, device="cuda")
    rays = 0
    for r in range(world):
        st, n_local = dist.render_tiles_device(scene, cam, r, world, packed[r])
        assert n_local == len(dist.local_tiles(w, h, r, world))
        rays += st.rays_total
    out = torch.empty(h * w * 3, dtype=torch.float32, device="cuda")
    dist.untile_device(packed, world, w, h, out)
    assert rays == st_full.rays_total
    np.testing.assert_allclose(out.cpu().numpy(), full.cpu().numpy(), rtol=0, atol=2e-6)
    np.testing.assert_allclose(dist.untile_host(packed.cpu().numpy(), world, w, h).reshape(-1), full.cpu().numpy(), atol=2e-6)
    scene.close()


@pytest.mark.parametrize("w,h,world", [(200, 120, 8), (203, 77, 3), (16, 16, 2)])
def test_tiles_resolved_straight_into_the_image(gpu, w, h, world):
    """The fused exchange (nrb_render_tiles_to_image): every virtual rank writes only its own pixels of ONE row-major
    image; after all ranks the image is the unsharded frame.  (203 wide: the scalar store path; ragged tiles.)"""
    import torch

    from nrays_b200 import dist

    scene, camd, cfg = configs.build("C3", target_tris=30000, lod=4)
    cam = make_camera(w, h, 2, 1.0, camd.eye, camd.projection((w, h)), seed=6)
    full = torch.empty(h * w * 3, dtype=torch.float32, device="cuda")
    dist.render_device(scene, cam, full)
    img = torch.full((h * w * 3,), -7.0, dtype=torch.float32, device="cuda")
    owned = np.zeros((h, w), bool)
    for r in range(world):
        before = img.clone()
        dist.render_tiles_to_image(scene, cam, r, world, img.data_ptr())
        changed = (img != before).reshape(h, w, 3).any(dim=2).cpu().numpy()
        mine = np.zeros((h, w), bool)
        tx, _ty = dist.tile_grid(w, h)
        for t in dist.local_tiles(w, h, r, world):
            y0, x0 = (t // tx) * 16, (t % tx) * 16
            mine[y0:y0 + 16, x0:x0 + 16] = True
        assert not (changed & ~mine).any(), "rank %d wrote pixels it does not own" % r
        owned |= mine
    assert owned.all()
    np.testing.assert_allclose(img.cpu().numpy(), full.cpu().numpy(), rtol=0, atol=2e-6)
    scene.close()


@pytest.mark.parametrize("w,h,world", [(192, 100, 4), (203, 77, 3)])
def test_rgb8_fused_into_the_tile_exchange(gpu, w, h, world):
    """SURVEY 8f-2 as written: RGB8 quantisation fused into the exchange.  Every (virtual) rank resolves its tiles straight into
    ONE row-major u8 image (device: peer-store form; host: 48-byte segments + strided 2-D DMA) — the bytes are the PNG
    quantisation (src/image.rs:64-77) of the unsharded float frame."""
    import mmap

    import torch

    from nrays_b200 import dist

    scene, camd, cfg = configs.build("C3", target_tris=30000, lod=4)
    cam = make_camera(w, h, 2, 1.0, camd.eye, camd.projection((w, h)), seed=9)
    full = np.empty(w * h * 3, np.float32)
    _lib.check(gpu.nrb_render(scene.handle, C.byref(cam), full.ctypes.data_as(C.POINTER(C.c_float)), None))
    exp = np.clip(full * np.float32(255.0), 0, 255).astype(np.uint8)
    img8 = torch.full((h * w * 3,), 77, dtype=torch.uint8, device="cuda")
    for r in range(world):
        dist.render_tiles_to_image_rgb8(scene, cam, r, world, img8.data_ptr())
    got = img8.cpu().numpy()
    assert (np.abs(exp.astype(int) - got.astype(int)) > 1).sum() == 0     # atomics order can flip a +-1 boundary
    assert (exp != got).mean() < 0.01
    if w % 16 == 0 and (w // 16) % world == 0:
        buf = mmap.mmap(-1, w * h * 3)
        host = np.frombuffer(buf, dtype=np.uint8)
        host[:] = 99
        cbuf = C.c_char.from_buffer(buf)
        dptr = C.c_void_p()
        _lib.check(gpu.nrb_host_register(0, C.c_void_p(C.addressof(cbuf)), w * h * 3, C.byref(dptr)))
        for r in range(world):
            dist.render_tiles_to_host_rgb8(scene, cam, r, world, C.addressof(cbuf))
        assert (np.abs(exp.astype(int) - host.astype(int)) > 1).sum() == 0
        assert (exp != host).mean() < 0.01
        _lib.check(gpu.nrb_host_unregister(0, C.c_void_p(C.addressof(cbuf))))
        del host, cbuf
        buf.close()
    else:
        ts = A.NrbTileSet(0, world)
        assert gpu.nrb_render_tiles_to_host_rgb8(scene.handle, C.byref(cam), C.byref(ts), C.c_void_p(img8.data_ptr()), None) == A.NRB_ERR_UNSUPPORTED
    scene.close()


# ---- driver paths that the default sizes never reach -------------------------------------------------
def _glass_mirror_scene():
    """Nodes that spawn BOTH children (refl_mix > 0 and alpha < 1): the tail kernel must spill one."""
    nodes = [node(Ball(0.9), phong(ka=(0.2, 0.1, 0.1)), pos=(-1.2, 0, 0), alpha=0.4, refr=1.2, refl=(0.3, 0.3)),
             node(Cuboid((0.7, 0.7, 0.7)), phong(ka=(0.1, 0.2, 0.1)), pos=(1.2, 0, 0.5), angle=(0, 30, 0), alpha=0.5, refr=1.1, refl=(0.25, 0.4)),
             node(Plane((0, 1, 0)), phong(), pos=(0, -1.0, 0), refl=(0.3, 0.3))]
    lights = [Light((1.0, 4.0, -2.0), 0.4, 10, (1, 1, 1))]   # 9 samples per hit
    return nodes, lights


def test_both_children_spill_and_area_light(gpu):
    nodes, lights = _glass_mirror_scene()
    img, st, ref, ost = render_both(nodes, lights, eye=(0, 1.5, -5.5), w=128, h=96, spp=2, window=1.0, seed=3)
    assert_parity(img, ref, what="both children", wh=(128, 96), twin=render_both.twin)
    assert_counts_close(st, ost)
    assert st.rays_reflect > 0 and st.rays_refract > 0


def test_tail_spill_bound_with_weak_attenuation(gpu):
    """Facing semi-transparent reflective panes with refl_atenuation 0.05: one chain reflects up to 19 times and spills a
    refraction ray at every bounce (the old 4x heuristic overflowed here).  The spill queue is now sized by the
    exact bound, and the frame matches the oracle with the tail phase forced on."""
    pane = phong(ka=(0.15, 0.2, 0.25), kd=(0.4, 0.4, 0.4))
    nodes = [node(Cuboid((1.6, 1.2, 0.05)), pane, pos=(0, 0, 1.0), alpha=0.5, refr=1.0, refl=(0.6, 0.05)),
             node(Cuboid((1.6, 1.2, 0.05)), pane, pos=(0, 0, -1.0), alpha=0.5, refr=1.0, refl=(0.6, 0.05)),
             node(Ball(0.4), phong(ka=(0.3, 0.1, 0.1)), pos=(0.2, 0.1, 0.0)),
             node(Plane((0, 1, 0)), phong(), pos=(0, -1.3, 0))]
    lights = [Light((0.5, 3.0, -4.0), 0.0, 1, (1, 1, 1))]
    for env in (dict(NRB_TAIL_RAYS=1 << 30, NRB_TAIL_MIN_WAVE=1), dict(NRB_TAIL_RAYS=1 << 30, NRB_TAIL_MIN_WAVE=1, NRB_SPILL_CAP=1000), dict()):
        with _Env(**env):
            img, st, ref, ost = render_both(nodes, lights, eye=(0.3, 0.4, -4.5), w=64, h=48, spp=1, window=0.0, max_depth=24)
        assert_parity(img, ref, what="weak attenuation panes %r" % (env,), wh=(64, 48), twin=render_both.twin)
        assert_counts_close(st, ost, rel=5e-3)
        assert st.rays_reflect > 10 * 64 * 48 / 4 and st.rays_refract > st.rays_reflect / 2


@pytest.mark.parametrize("env", [dict(NRB_TAIL_RAYS=0), dict(NRB_TAIL_RAYS=1 << 30), dict(NRB_SHADOW_CAP=4096),
                                 dict(NRB_BATCH_SLOTS=4096), dict(NRB_BATCH_SLOTS=4096, NRB_SHADOW_CAP=2048, NRB_TAIL_RAYS=64),
                                 dict(NRB_REVERSE_SHADOW=0), dict(NRB_NODE_FORMAT=2), dict(NRB_NODE_FORMAT=2, NRB_REFILL_RAYS=24, NRB_REFILL_SHADOW=24),
                                 dict(NRB_NODE_FORMAT=3), dict(NRB_NODE_FORMAT=4, NRB_REFILL_RAYS=24, NRB_REFILL_SHADOW=24),
                                 dict(NRB_REFILL_PRIMARY=20, NRB_REFILL_RAYS=24, NRB_REFILL_SHADOW=24),
                                 dict(NRB_REFILL_PRIMARY=31, NRB_REFILL_RAYS=31, NRB_REFILL_SHADOW=31, NRB_TAIL_RAYS=0)])
def test_driver_paths_give_the_same_image(gpu, env):
    """No tail / everything in the tail / chunked shadow queue / many batches / dynamic fetch (warps refill their
    idle lanes mid-packet): same frame as the default path."""
    nodes, lights = _glass_mirror_scene()
    base, st0, ref, ost = render_both(nodes, lights, eye=(0, 1.5, -5.5), w=96, h=80, spp=2, window=1.0, seed=8)
    with _Env(**env):
        img, st, _, _ = render_both(nodes, lights, eye=(0, 1.5, -5.5), w=96, h=80, spp=2, window=1.0, seed=8)
    np.testing.assert_allclose(img, base, rtol=0, atol=3e-5)   # only the order of float atomics may differ
    assert st.as_dict()["rays_total"] == st0.as_dict()["rays_total"]
    assert_parity(img, ref, what=str(env), wh=(96, 80), twin=render_both.twin)


@pytest.mark.timeout(180)
def test_node_formats_and_speculative_loop_on_small_scenes(gpu):
    """The node formats picked for large scenes — 2 (bf16 half extents, speculative while-while loop), 3 (16-bit grid cells,
    32-byte records) and 4 (grid + speculative loop) — forced on scenes whose traversal STARTS on a leaf (one shape, <= 4
    triangles), on the shape zoo and on meshes: same image as format 0.  The grid formats exist for mesh-only scenes; the
    library keeps format 0 elsewhere, which the build info reports."""
    P1 = np.array([[-1, -1, 0], [1, -1, 0], [0, 1, 0]], np.float32)
    scenes = [
        ([node(Ball(1.0), NormalMaterial())], [], False),
        ([node(TriMesh(P1, np.array([[0, 1, 2]], np.uint32), None), NormalMaterial())], [], True),
        ([node(TriMesh(*quad_mesh(1.0, 1)), phong()), node(Plane((0, 1, 0)), phong(), pos=(0, -1, 0))], [Light((1, 4, -2), 0.0, 1, (1, 1, 1))], False),
        (zoo([phong(), phong(), phong(), phong(), phong()]), [Light((2, 4, -3), 0.0, 1, (1, 1, 1))], False),
        ([node(TriMesh(*quad_mesh(1.5, 6, y=0.3)), phong(), alpha=0.5, refr=1.2), node(TriMesh(*quad_mesh(3.0, 5, y=-0.5)), phong(), refl=(0.3, 0.4))],
         [Light((0.5, 5, -1), 0.3, 4, (1, 1, 1))], True),
        ([node(TriMesh(*quad_mesh(2.0, 24, y=0.0)), phong(), refl=(0.2, 0.3)), node(TriMesh(*quad_mesh(0.7, 9, y=0.8)), phong(), pos=(0.2, 0, 0.3)),
          node(TriMesh(*quad_mesh(200.0, 3, y=-0.6)), phong())], [Light((0.5, 5, -1), 0.0, 1, (1, 1, 1)), Light((-3, 2, -2), 0.2, 2, (0.5, 0.5, 0.6))], True),
    ]
    for k, (nodes, lights, mesh_only) in enumerate(scenes):
        base = None
        for fmt in (0, 2, 3, 4):
            with _Env(NRB_NODE_FORMAT=fmt):
                img, st, _, _ = render_both(nodes, lights, eye=(0.3, 1.5, -5.0), w=80, h=60, spp=2, window=1.0, seed=k)
                used = Scene(nodes, lights, (1.0, 1.0, 1.0)).build_info().node_format
            assert used == (fmt if ((mesh_only and k != 1) or fmt in (0, 2)) else 0), (k, fmt, used)   # scene 1 has no inner node: no grid
            if base is None:
                base = (img, st)
                continue
            d = np.abs(base[0] - img).max(axis=1)
            assert (d > 1e-4).mean() < 5e-3, (k, fmt, float((d > 1e-4).mean()))   # identical hits up to exact ties / box rounding
            assert abs(int(base[1].rays_total) - int(st.rays_total)) <= max(4, 2e-3 * base[1].rays_total), (k, fmt)


def test_mesh_scene_driver_paths(gpu):
    scene, camd, cfg = configs.build("C3", target_tris=30000, lod=4)
    w, h = 160, 90
    cam = make_camera(w, h, 2, 1.0, camd.eye, camd.projection((w, h)), seed=2)

    def go():
        out = np.empty((w * h, 3), np.float32)
        st = A.NrbStats()
        _lib.check(gpu.nrb_render(scene.handle, C.byref(cam), out.ctypes.data_as(C.POINTER(C.c_float)), C.byref(st)))
        return out, st

    base, st0 = go()
    for env in (dict(NRB_TAIL_RAYS=0), dict(NRB_TAIL_RAYS=1 << 30), dict(NRB_BATCH_SLOTS=8192), dict(NRB_SHADOW_CAP=1024),
                dict(NRB_REFILL_PRIMARY=16, NRB_REFILL_RAYS=20, NRB_REFILL_SHADOW=20), dict(NRB_REVERSE_SHADOW=0)):
        with _Env(**env):
            img, st = go()
        np.testing.assert_allclose(img, base, rtol=0, atol=3e-5, err_msg=str(env))
        assert st.as_dict()["rays_total"] == st0.as_dict()["rays_total"], env
    scene.close()


def test_grid_format_needs_a_grid_that_resolves_the_geometry(gpu):
    """One far-away triangle stretches the 16-bit grid until a cell is larger than the mesh that matters: the library keeps
    float boxes (format 4 requested -> 2, format 3 -> 0) instead of letting every snapped box overlap its neighbours."""
    P, I, UV = quad_mesh(1.0, 48, y=0.0)
    far = np.array([[4.0e5, 0, 4.0e5], [4.0e5 + 1, 0, 4.0e5], [4.0e5, 0, 4.0e5 + 1]], np.float32)
    P2 = np.concatenate([P, far]).astype(np.float32)
    I2 = np.concatenate([I, np.array([[len(P), len(P) + 1, len(P) + 2]], np.uint32)]).astype(np.uint32)
    nodes = [node(TriMesh(P2, I2, None), phong())]
    lights = [Light((0.5, 5, -1), 0.0, 1, (1, 1, 1))]
    imgs = {}
    for fmt, want in ((0, 0), (3, 0), (4, 2)):
        with _Env(NRB_NODE_FORMAT=fmt):
            img, st, _, _ = render_both(nodes, lights, eye=(0.3, 1.5, -5.0), w=64, h=48, spp=1, window=0.0, seed=0)
            assert Scene(nodes, lights, (1.0, 1.0, 1.0)).build_info().node_format == want, fmt
        imgs[fmt] = img
    np.testing.assert_allclose(imgs[3], imgs[0], rtol=0, atol=3e-5)
    np.testing.assert_allclose(imgs[4], imgs[0], rtol=0, atol=3e-5)
    # ... and with the limit lifted the grid format still renders the same frame (correct, only slow on real scenes)
    with _Env(NRB_NODE_FORMAT=3, NRB_GRID_MAX_INFLATION=1e30):
        img, st, _, _ = render_both(nodes, lights, eye=(0.3, 1.5, -5.0), w=64, h=48, spp=1, window=0.0, seed=0)
        assert Scene(nodes, lights, (1.0, 1.0, 1.0)).build_info().node_format == 3
    np.testing.assert_allclose(img, imgs[0], rtol=0, atol=3e-5)


def test_grid_format_far_from_the_origin(gpu):
    """A mesh 2000 units from the origin, 2 units wide: the grid cell is floored at four float steps of the coordinates, so the
    one cell of margin still covers the f32 rounding of the slab test: same frame and ray counts as format 0."""
    off = np.array([1000.0, 500.0, -2000.0])
    nodes = [node(TriMesh(*quad_mesh(1.0, 24, y=0.0)), phong(), pos=tuple(off), refl=(0.3, 0.4)),
             node(TriMesh(*quad_mesh(0.4, 6, y=0.5)), phong(), pos=tuple(off + (0.1, 0.0, 0.2)))]
    lights = [Light(tuple(off + (0.5, 5.0, -1.0)), 0.0, 1, (1, 1, 1))]
    eye = tuple(off + (0.3, 1.5, -5.0))
    imgs = {}
    for fmt in (0, 3, 4):
        with _Env(NRB_NODE_FORMAT=fmt):
            img, st, ref, ost = render_both(nodes, lights, eye=eye, at=tuple(off), w=96, h=64, spp=2, window=1.0, seed=4)
            assert Scene(nodes, lights, (1.0, 1.0, 1.0)).build_info().node_format == fmt
        imgs[fmt] = (img, st)
    assert imgs[0][1].rays_shadow > 0 and imgs[0][1].rays_reflect > 0
    for fmt in (3, 4):
        np.testing.assert_allclose(imgs[fmt][0], imgs[0][0], rtol=0, atol=3e-5)
        assert imgs[fmt][1].rays_total == imgs[0][1].rays_total


@pytest.mark.parametrize("cfg_id,kw", [("C3", dict(target_tris=30000, lod=4)), ("C4", dict(target_tris=40000))])
def test_node_formats_on_meshes(gpu, cfg_id, kw):
    """Every node format renders the mesh configs to the same frame (identical hits: the device boxes only ever grow)."""
    w, h = 160, 90
    base = None
    for fmt in (0, 2, 3, 4):
        with _Env(NRB_NODE_FORMAT=fmt):
            scene, camd, cfg = configs.build(cfg_id, **kw)
        assert scene.build_info().node_format == fmt
        cam = make_camera(w, h, 2, 1.0, camd.eye, camd.projection((w, h)), seed=2)
        out = np.empty((w * h, 3), np.float32)
        st = A.NrbStats()
        _lib.check(gpu.nrb_render(scene.handle, C.byref(cam), out.ctypes.data_as(C.POINTER(C.c_float)), C.byref(st)))
        scene.close()
        if base is None:
            base = (out, st)
            continue
        np.testing.assert_allclose(out, base[0], rtol=0, atol=3e-5, err_msg="format %d" % fmt)
        assert st.as_dict()["rays_total"] == base[1].as_dict()["rays_total"], fmt


@pytest.mark.parametrize("builder,code", [("lbvh", 1), ("ploc", 2)])
def test_device_bvh_builds_render_the_same_image(gpu, builder, code):
    """§8f-1: the GPU builds (Morton sort + Karras radix tree, or + PLOC clustering; bottom-up fit, collapse to <= 4-triangle
    leaves) are different trees over the same triangles: same pixels as the host SAH build, valid structure, built on the device."""
    from nrays_b200.loader3d import load_scene

    cfg = configs.CONFIGS["C3"]
    with _Env(NRB_CHECK_BVH=1):
        text, res = cfg["text"](), cfg["resolver"](target_tris=60000, lod=4)
        lights, nodes, cams = __import__("nrays_b200.loader3d", fromlist=["parse"]).parse(text, res)
        sah = Scene(nodes, lights, (1, 1, 1), builder="sah")
        dev = Scene(nodes, lights, (1, 1, 1), builder=builder)
    bi_s, bi_l = sah.build_info(), dev.build_info()
    assert bi_s.builder == A.NRB_BUILDER_SAH and bi_l.builder == code
    assert bi_l.gpu_build_ms > 0 and bi_s.gpu_build_ms == 0
    assert bi_l.triangles == bi_s.triangles == 60000 and bi_l.max_depth <= 60
    w, h = 192, 108
    proj = cams[0].projection((w, h))
    a, sa = render(sah, (w, h), 2, 1.0, cams[0].eye, proj, seed=4, return_stats=True)
    b, sb = render(dev, (w, h), 2, 1.0, cams[0].eye, proj, seed=4, return_stats=True)
    # identical hits except exact ties (shared edges), so only a handful of pixels may move
    d = np.abs(a.pixels - b.pixels).max(axis=1)
    assert (d > 1e-4).mean() < 2e-3, float((d > 1e-4).mean())
    assert abs(int(sa.rays_total) - int(sb.rays_total)) <= 1e-3 * sa.rays_total
    sah.close()
    dev.close()
    # hair: thin, overlapping geometry and many duplicate Morton cells
    scene, camd, _ = configs.build("C4", target_tris=160000)
    hair_nodes, hair_lights = scene.nodes, scene.lights()
    with _Env(NRB_CHECK_BVH=1):
        lb = Scene(hair_nodes, hair_lights, (1, 1, 1), builder=builder)
    proj = camd.projection((128, 72))
    a = render(scene, (128, 72), 1, 0.0, camd.eye, proj)
    b = render(lb, (128, 72), 1, 0.0, camd.eye, proj)
    assert (np.abs(a.pixels - b.pixels).max(axis=1) > 1e-4).mean() < 2e-3
    scene.close()
    lb.close()
    # tiny and degenerate inputs: <= 4 triangles (a single leaf), 5 triangles, all triangles identical (one Morton cell)
    for n_tri, same in ((3, False), (5, False), (40, True)):
        rng = np.random.default_rng(n_tri)
        base = rng.uniform(-1, 1, (1, 3, 3)).astype(np.float32)
        P = (np.repeat(base, n_tri, 0) if same else rng.uniform(-1, 1, (n_tri, 3, 3)).astype(np.float32)).reshape(-1, 3)
        Fc = np.arange(3 * n_tri, dtype=np.uint32).reshape(-1, 3)
        nd = [node(TriMesh(P, Fc, None), NormalMaterial())]
        with _Env(NRB_CHECK_BVH=1):
            s1, s2 = Scene(nd, [], (0, 0, 0), builder="sah"), Scene(nd, [], (0, 0, 0), builder=builder)
        pr = camera_projection((0, 0, -4), (0, 0, 0), 45.0, 48, 48)
        i1, i2 = render(s1, (48, 48), 1, 0.0, (0, 0, -4), pr), render(s2, (48, 48), 1, 0.0, (0, 0, -4), pr)
        assert (np.abs(i1.pixels - i2.pixels).max(axis=1) > 1e-4).mean() < 5e-3, (n_tri, same)
        s1.close()
        s2.close()
