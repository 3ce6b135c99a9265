"""The bench.py contract the driver depends on: ONE JSON line on stdout with the agreed keys.

* reference arm (`--impl reference`): runs here on the CPU (the oracle port timed on the host cores; the one place outside
  tests/ and smoke() where oracle/ is executed) — checked in the CPU suite;
* native arm: a short run on the GPU (`-m gpu`): roofline / e2e / gpu_launches / clocks present and sane, and the library's
  own kernels did the work (gpu_launches > 0).
"""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e"}


def _run(*args, timeout=900):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout,
                       cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, f"stdout must carry exactly one line, got {len(lines)}: {r.stdout[:500]}"
    return json.loads(lines[0])


def test_reference_arm_prints_the_contract_line():
    line = _run("--impl", "reference", "--steps", "1", "--warmup", "1")
    assert BASE_KEYS <= set(line), BASE_KEYS - set(line)
    assert line["impl"] == "reference" and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["unit"] == "samples/s" and line["value"] > 0 and line["dtype"] == "fp32" and line["data"] == "synthetic"
    assert "workload" in line["config"] and "model" not in line["config"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    e = line["e2e"]
    assert e["value"] == line["value"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


@pytest.mark.gpu
def test_native_arm_prints_the_contract_line(cuda_lib):
    line = _run("--steps", "4", "--warmup", "3", "--no-cpu", "--no-secondary", "--no-update")
    assert BASE_KEYS <= set(line), BASE_KEYS - set(line)
    assert "impl" not in line or line["impl"] != "reference"
    assert line["n_gpus"] == 1 and line["steps"] == 4 and line["scaling"] == "weak" and line["dtype"] == "bf16"
    assert line["value"] > 1e4 and line["gpu_launches"] > 100
    assert "workload" in line["config"] and "model" not in line["config"]
    e = line["e2e"]
    assert 0 < e["value"] <= 1.05 * line["value"] and e["h2d_bytes_per_step"] > 4e8 and e["d2h_bytes_per_step"] > 0
    r = line["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s") and r["peak"] > 0
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-6 and 0.2 < r["frac"] < 1.0
    assert r["traffic"] is None or r["traffic"] > 0
    c = line["clocks"]
    assert c["sm_mhz"] > 0 and c["sm_max_mhz"] >= c["sm_mhz"] and isinstance(c["reasons"], list)
    assert not ({"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(c["reasons"]))
