"""Generates the committed golden fixtures from the f64 CPU oracle.

The reference has no golden images or known-answer vectors and cannot be run here (SURVEY.md §4,
§8c), so these fixtures pin the ORACLE (regression across refactors) and give the GPU tests a
device-independent target; they are not reference outputs.  Regenerate with
    python tests/golden/make_golden.py
and commit the .npz files together with this script.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle_lib as O  # noqa: E402
from nrays_b200 import configs, make_camera  # noqa: E402
from nrays_b200.loader3d import load_scene  # noqa: E402

# name -> (config, build kwargs, scene-text kwargs, w, h, spp, window, seed)
CASES = {
    "c1_rngfree": ("C1", dict(globe_size=(64, 32)), dict(light_radius=0.0), 64, 64, 1, 0.0, 0),
    "c1_asis": ("C1", dict(globe_size=(64, 32)), dict(), 64, 64, 1, 0.0, 0),
    "c1_aa": ("C1", dict(globe_size=(64, 32)), dict(), 48, 48, 4, 1.0, 3),
    "c2": ("C2", dict(globe_size=(256, 128)), dict(), 64, 64, 2, 1.0, 0),
    "c3_lod4": ("C3", dict(target_tris=30000, lod=4), dict(), 96, 54, 1, 0.0, 0),
    "c4_small": ("C4", dict(target_tris=80000), dict(), 64, 36, 2, 1.0, 0),
}


def build_case(name, upload=False, device=0):
    cfgname, kw, textkw, w, h, spp, window, seed = CASES[name]
    cfg = configs.CONFIGS[cfgname]
    scene, cameras = load_scene(cfg["text"](**textkw), cfg["resolver"](**kw), device=device, upload=upload)
    camd = cameras[0]
    cam = make_camera(w, h, spp, window, camd.eye, camd.projection((w, h)), seed=seed)
    return scene, camd, cam, (w, h, spp, window, seed)


def main():
    for name in CASES:
        scene, camd, cam, (w, h, spp, window, seed) = build_case(name)
        img, st = O.OracleScene(scene.flat, 64).render(cam)
        counts = np.array([st.rays_primary, st.rays_reflect, st.rays_refract, st.rays_shadow, st.paths_truncated], np.int64)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), image=img.astype(np.float32), counts=counts,
                            shape=np.array([w, h, spp], np.int64))
        print(name, img.shape, counts.tolist())


if __name__ == "__main__":
    main()
