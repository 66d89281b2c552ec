"""bench.py keeps the driver's JSON contract (one line, required keys) for both arms."""
import json
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches"}


def _run(args, timeout=900):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                       timeout=timeout, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_json_contract():
    d = _run(["--impl", "reference", "--steps", "1", "--warmup", "1", "--height", "64", "--width", "96"])
    assert BASE_KEYS <= set(d)
    assert d["impl"] == "reference" and d["unit"] == "pairs/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert "workload" in d["config"] and "model" not in d["config"]


@pytest.mark.gpu
def test_ours_json_contract_on_gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    d = _run(["--steps", "5", "--warmup", "3", "--e2e-steps", "4", "--train-steps", "0", "--cpu-reps", "1"])
    assert BASE_KEYS | {"roofline", "cpu_baseline", "clocks"} <= set(d)
    assert "impl" not in d and d["dtype"] == "f32" and d["scaling"] == "weak" and d["n_gpus"] == 1
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["peak"] > 1000
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0 < r["frac"] < 1.5
    # the dominant kernel is timed twice: alone inside the step, and in place inside the step's CUDA graph
    assert r["us_per_launch"] > 0 and r["us_per_launch_in_place"] >= 0.9 * r["us_per_launch"] and r["timing"]
    assert abs(r["achieved"] - r["alg_bytes_per_launch"] / r["us_per_launch"] / 1e3) < 1e-6 * r["achieved"]
    assert "one CUDA graph" in d["config"]["step"]
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["e2e"]["value"] < d["value"]
    assert d["gpu_launches"] >= 5 * 60                      # >= 60 of our kernels per step
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] > 0
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
