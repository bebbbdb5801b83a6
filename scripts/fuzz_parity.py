"""Randomised parity sweep (developer tool, GPU): random scenes mixing every shape, meshes with and without opacity
maps, transparent / reflective / refractive nodes, point and area lights, random cameras and sample counts, rendered
on the device through the C-ABI under random driver / kernel knobs and compared with the CPU oracle.

    python scripts/fuzz_parity.py [n_scenes] [seed]

Prints one line per scene and a summary; exits 1 if any scene is outside the stated tolerance
(per-channel |delta| <= 1/255 on >= 99.5 % of the pixels of these tiny images, ray counts within 0.5 %) without a proof that it
is precision-chaotic (the same proof tests/test_fuzz_gpu.py demands of its fixed-seed slice).
A scene that fails against the f64 oracle is re-checked against the oracle's f32 mode.  Scenes with deep mirror + glass
recursion (max_depth 12, both children at every hit) have chaotic path trees: one branch flipped by rounding changes the
ray count by thousands while the image stays within tolerance — the f64 and f32 oracles disagree with each other
on those too.  Round 1: seed 1, 150 scenes: 145 clean, 5 of that kind; seed 7 (with depth-shift nodes), 200 scenes: 195 clean,
5 of that kind (the f64 and f32 oracles differ from each other by 0.1-4.5 % of the pixels on them).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from nrays_b200 import (Ball, Capsule, Cone, Cuboid, Cylinder, ImageData, Interpolation, Light, NormalMaterial, Overflow,  # noqa: E402
                        PhongMaterial, Plane, Texture2d, TriMesh, UVMaterial)
from util import image_metrics, node, quad_mesh, render_both  # noqa: E402

KNOBS = [{}, {"NRB_TAIL_RAYS": "0"}, {"NRB_TAIL_RAYS": str(1 << 30)}, {"NRB_REVERSE_SHADOW": "0"},
         {"NRB_REFILL_PRIMARY": "24", "NRB_REFILL_RAYS": "28", "NRB_REFILL_SHADOW": "28"},
         {"NRB_REFILL_PRIMARY": "8", "NRB_REFILL_RAYS": "12", "NRB_REFILL_SHADOW": "12", "NRB_SMALL_QUEUE": "0"},
         {"NRB_BATCH_SLOTS": "2048", "NRB_SHADOW_CAP": "1024"}, {"NRB_SMALL_QUEUE": str(0xFFFFFFFF)}]


def rand_texture(rng, alpha=False):
    w, h = int(rng.integers(2, 12)), int(rng.integers(2, 12))
    px = rng.uniform(0.0, 1.0, (h * w, 4)).astype(np.float32)
    if alpha:
        px[:, :3] = 1.0
        px[:, 3] = (rng.uniform(size=h * w) > 0.45).astype(np.float32) if rng.uniform() < 0.6 else rng.uniform(size=h * w)
    else:
        px[:, 3] = 1.0
    interp = Interpolation.Bilinear if rng.uniform() < 0.7 else Interpolation.Nearest
    over = Overflow.Wrap if rng.uniform() < 0.7 else Overflow.ClampToEdges
    return Texture2d(ImageData(px, (w, h)), interp, over)


def rand_material(rng, allow_alpha_map):
    r = rng.uniform()
    if r < 0.1:
        return NormalMaterial()
    if r < 0.2:
        return UVMaterial()
    tex = rand_texture(rng) if rng.uniform() < 0.4 else None
    amap = rand_texture(rng, alpha=True) if (allow_alpha_map and rng.uniform() < 0.35) else None
    return PhongMaterial(tuple(rng.uniform(0.0, 0.4, 3)), tuple(rng.uniform(0.2, 1.0, 3)), tuple(rng.uniform(0.0, 1.0, 3)), tex, amap,
                         float(rng.choice([1.0, 8.0, 40.0, 150.0])))


def rand_scene(rng):
    nodes = []
    n = int(rng.integers(1, 8))
    for _ in range(n):
        k = int(rng.integers(0, 7))
        if k == 0:
            g = Ball(float(rng.uniform(0.3, 1.2)))
        elif k == 1:
            g = Cuboid(tuple(rng.uniform(0.2, 1.0, 3)))
        elif k == 2:
            g = Cylinder(float(rng.uniform(0.3, 1.0)), float(rng.uniform(0.2, 0.8)))
        elif k == 3:
            g = Capsule(float(rng.uniform(0.2, 0.8)), float(rng.uniform(0.2, 0.6)))
        elif k == 4:
            g = Cone(float(rng.uniform(0.3, 1.0)), float(rng.uniform(0.2, 0.8)))
        else:
            P, F, UV = quad_mesh(float(rng.uniform(0.5, 2.5)), int(rng.integers(1, 7)), y=0.0)
            P = P + rng.normal(0.0, 0.05, P.shape).astype(np.float32)   # not axis-aligned
            g = TriMesh(P, F, UV if rng.uniform() < 0.85 else None)
        mesh = k >= 5
        alpha = 1.0 if rng.uniform() < 0.6 else float(rng.uniform(0.1, 0.9))
        refl = (0.0, 0.0) if rng.uniform() < 0.6 else (float(rng.uniform(0.1, 0.9)), float(rng.choice([0.2, 0.35, 0.5])))
        refr = 1.0 if rng.uniform() < 0.5 else float(rng.uniform(1.05, 1.8))
        nd = node(g, rand_material(rng, mesh), pos=tuple(rng.uniform(-2.5, 2.5, 3)), angle=tuple(rng.uniform(-180, 180, 3)),
                  refl=refl, alpha=alpha, refr=refr, solid=bool(rng.uniform() < 0.15))
        if rng.uniform() < 0.12:   # SceneNode.nmap: depth-shift texture (general trace kernel)
            t = rand_texture(rng)
            t.data.pixels[:, :3] *= float(rng.uniform(0.2, 1.0))
            nd.nmap = t
        nodes.append(nd)
    if rng.uniform() < 0.6:
        nodes.append(node(Plane((0, 1, 0)), rand_material(rng, False), pos=(0, float(rng.uniform(-3.5, -2.0)), 0),
                          refl=(0.0, 0.0) if rng.uniform() < 0.5 else (0.3, 0.5), alpha=1.0 if rng.uniform() < 0.8 else 0.5))
    lights = []
    for _ in range(int(rng.integers(0, 3))):
        radius = 0.0 if rng.uniform() < 0.6 else float(rng.uniform(0.05, 0.5))
        lights.append(Light(tuple(rng.uniform(-5, 5, 3) + np.array([0, 5, 0])), radius, int(rng.choice([1, 4, 9])) if radius else 1,
                            tuple(rng.uniform(0.3, 1.0, 3))))
    return nodes, lights


def main():
    n_scenes = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rng = np.random.default_rng(seed)
    bad = 0
    for i in range(n_scenes):
        nodes, lights = rand_scene(rng)
        eye = tuple(rng.uniform(-1, 1, 3) * 2.0 + np.array([0.0, 1.0, -7.0]))
        w, h = int(rng.integers(17, 80)), int(rng.integers(9, 60))
        spp = int(rng.integers(1, 4))
        window = 0.0 if rng.uniform() < 0.4 else 1.0
        knobs = KNOBS[int(rng.integers(0, len(KNOBS)))]
        os.environ.update(knobs)
        try:
            img, st, ref, ost = render_both(nodes, lights, eye=eye, w=w, h=h, spp=spp, window=window, seed=i, max_depth=12)
        finally:
            for k in knobs:
                os.environ.pop(k, None)
        m = image_metrics(img, ref)
        finite = bool(np.isfinite(img).all())
        counts_ok = all(abs(int(getattr(st, k)) - int(getattr(ost, k))) <= max(6, 5e-3 * int(getattr(ost, k)))
                        for k in ("rays_primary", "rays_reflect", "rays_refract", "rays_shadow", "paths_truncated"))
        ok = finite and m["frac_over"] <= 5e-3 and counts_ok
        note = ""
        if not ok:
            # precision or logic?  The f32 mode of the oracle is the device's arithmetic twin: a scene that is off against
            # the f64 oracle but agrees with the twin is a chaotic path tree (deep mirror / glass recursion), not a bug.
            from nrays_b200 import Scene, make_camera
            from util import look
            import oracle_lib as O
            sc = Scene(nodes, lights, (1.0, 1.0, 1.0), upload=False)
            cam = make_camera(w, h, spp, window, eye, look(eye, (0, 0, 0), 45.0, w, h), seed=i, max_depth=12)
            ref32, ost32 = O.OracleScene(sc.flat, 32).render(cam)
            m32 = image_metrics(img, ref32)
            c32 = all(abs(int(getattr(st, k)) - int(getattr(ost32, k))) <= max(6, 5e-3 * int(getattr(ost32, k)))
                      for k in ("rays_primary", "rays_reflect", "rays_refract", "rays_shadow", "paths_truncated"))
            m6432 = image_metrics(ref, ref32)
            note = " | vs f32 twin: frac_over=%.4f rays=%d counts_%s | f64 vs f32 oracle: frac_over=%.4f" % (
                m32["frac_over"], ost32.rays_reference, "ok" if c32 else "DIFFER", m6432["frac_over"])
            keys = ("rays_primary", "rays_reflect", "rays_refract", "rays_shadow", "paths_truncated")
            note += " | classes dev/f64/f32: " + " ".join("%s=%d/%d/%d" % (k[5:] if k.startswith("rays_") else k, getattr(st, k), getattr(ost, k),
                                                                            getattr(ost32, k)) for k in keys)
            if m32["frac_over"] <= 5e-3 and c32:
                ok = True
                note += " -> precision-chaotic, agrees with the twin"
            else:
                # the proof tests/test_fuzz_gpu.py demands: every over-tolerance pixel that is not an edge flip lies where the
                # oracle's own f64 and f32 modes disagree (or the device equals the twin), and the ray counts agree with the
                # twin or the two oracle modes disagree with EACH OTHER by more than the tolerance
                from util import TOL, _nbhd_min_max
                a = np.asarray(img, np.float64).reshape(h, w, 3)
                b = np.asarray(ref, np.float64).reshape(h, w, 3)
                c = np.asarray(ref32, np.float64).reshape(h, w, 3)
                lo, hi = _nbhd_min_max(b)
                over = np.abs(a - b).max(axis=2) > TOL
                flip = over & ((((hi - lo).max(axis=2) > 2 * TOL) & ((a >= lo - TOL) & (a <= hi + TOL)).all(axis=2)) | ((hi - lo).max(axis=2) > 0.1))
                unproven = over & ~flip & ~((np.abs(b - c).max(axis=2) > TOL) | (np.abs(a - c).max(axis=2) <= TOL))
                oracles_differ = not all(abs(int(getattr(ost32, k)) - int(getattr(ost, k))) <= max(6, 5e-3 * int(getattr(ost, k))) for k in keys)
                if not unproven.any() and (c32 or oracles_differ):
                    ok = True
                    note += " -> precision-chaotic: 0 unproven pixels, the f64 and f32 oracles disagree with each other"
                else:
                    note += " -> %d unproven pixels, oracle modes %s on the counts" % (int(unproven.sum()), "disagree" if oracles_differ else "AGREE")
        bad += 0 if ok else 1
        print("%3d %s %dx%dx%d nodes=%d lights=%d knobs=%s frac_over=%.4f max=%.3f rays=%d/%d culled=%d %s" % (
            i, "ok " if ok else "BAD", w, h, spp, len(nodes), len(lights), ",".join(knobs) or "-", m["frac_over"], m["max_abs"],
            st.rays_reference, ost.rays_reference, st.rays_shadow_culled, ("" if counts_ok else "COUNTS") + note), flush=True)
    print("fuzz: %d scenes, %d outside tolerance" % (n_scenes, bad))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
