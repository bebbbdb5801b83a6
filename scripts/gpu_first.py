"""First GPU contact: render configs on the device and on the oracle, report the parity metrics.
Writes images + a JSON report under gpurun_out/."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O  # noqa: E402
from nrays_b200 import configs, make_camera, render  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)


def save_png(name, img, w, h):
    from PIL import Image

    Image.fromarray((np.clip(img.reshape(h, w, 3), 0, 1) * 255).astype(np.uint8)).save(os.path.join(OUT, name))


def compare(tag, scene, camdesc, w, h, spp, window, seed=0, bits=64, **kw):
    proj = camdesc.projection((w, h))
    t0 = time.time()
    img, st = render(scene, (w, h), spp, window, camdesc.eye, proj, seed=seed, return_stats=True)
    t_gpu = time.time() - t0
    cam = make_camera(w, h, spp, window, camdesc.eye, proj, seed=seed)
    osc = O.OracleScene(scene.flat, bits)
    t0 = time.time()
    ref, ost = osc.render(cam)
    t_cpu = time.time() - t0
    d = np.abs(img.pixels - ref).max(axis=1)
    rep = dict(tag=tag, w=w, h=h, spp=spp, max_abs=float(d.max()), mean_abs=float(np.abs(img.pixels - ref).mean()),
               frac_over=float((d > 1.0 / 255.0).mean()), gpu=st.as_dict(), oracle=ost.as_dict(), t_gpu_s=t_gpu,
               t_cpu_s=t_cpu)
    save_png(tag + "_gpu.png", img.pixels, w, h)
    save_png(tag + "_ref.png", ref, w, h)
    save_png(tag + "_diff.png", np.minimum(np.abs(img.pixels - ref) * 20.0, 1.0), w, h)
    print(json.dumps(rep))
    return rep


def main():
    reps = []
    scene, cam, cfg = configs.build("C1")
    reps.append(compare("c1_asis", scene, cam, 256, 256, 1, 0.0))
    reps.append(compare("c1_aa", scene, cam, 128, 128, 4, 1.0))
    scene, cam, cfg = configs.build("C2", globe_size=(512, 256))
    reps.append(compare("c2", scene, cam, 256, 256, 4, 1.0))
    scene, cam, cfg = configs.build("C3", target_tris=40000, lod=4)
    reps.append(compare("c3_small", scene, cam, 320, 180, 1, 0.0))
    scene, cam, cfg = configs.build("C4", target_tris=160000)
    reps.append(compare("c4_small", scene, cam, 256, 144, 1, 0.0))
    json.dump(reps, open(os.path.join(OUT, "gpu_first.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
