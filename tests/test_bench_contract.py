"""bench.py's JSON contract: the reference arm runs on CPU here (config C1, the reference's own CPU-runnable case); the GPU arm is
checked on the B200 box."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
             "data", "config", "e2e", "cpu_baseline", "gpu_launches"}


def _run(*args, timeout=600):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line_on_cpu():
    line = _run("--impl", "reference", "--config", "C1", "--steps", "2", "--warmup", "1")
    assert BASE_KEYS <= set(line), BASE_KEYS - set(line)
    assert line["impl"] == "reference" and line["unit"] == "Mrays/s" and line["higher_is_better"] is True
    assert line["steps"] == 2 and line["warmup"] == 1 and line["value"] > 0 and line["ms_per_step"] > 0
    assert line["config"]["workload"].startswith("primitives.scene") and "sample" not in line["config"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == os.cpu_count() and cb["value"] == line["value"] and "full frame" in cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["vs_baseline"] is None and line["gpu_launches"] == 0


def test_reference_arm_only_rank_zero_prints():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "C1", "--gpus", "2", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


@pytest.mark.gpu
def test_gpu_arm_line(gpu):
    line = _run("--config", "C1", "--steps", "3", "--warmup", "3", "--extra", "C2", "--no-cpu-baseline")
    assert BASE_KEYS <= set(line) and {"roofline", "clocks", "configs"} <= set(line)
    assert line["n_gpus"] == 1 and line["gpu_launches"] > 0 and line["dtype"] == "f32"
    r = line["roofline"]
    assert r["kernel"] == "trace_kernel" and r["bound"] == "hbm" and 0 < r["frac"] < 1.5 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert {"frame", "other_kernels_ms_per_step", "kernel_share_of_step", "launches_per_step"} <= set(r)
    e = line["e2e"]
    assert e["d2h_bytes_per_step"] == 256 * 256 * 3 * 4 and e["h2d_bytes_per_step"] > 0 and e["value"] > 0 and e["ms_per_step"] > 0
    assert "C2" in line["configs"] and line["configs"]["C2"]["value"] > 0 and "roofline" in line["configs"]["C2"]
    assert line["config"]["workload"].startswith("primitives.scene")
