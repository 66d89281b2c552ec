"""The flow_warp kernels replace the two IEEE divisions of the reference's coordinate pipeline
(utils/warp_utils.py:16-23: 2*v/(size-1)) by a reciprocal multiply plus one FMA correction
(unsamflow_b200/csrc/warp_coord.cuh: div_by_size).  That is only a drop-in if the result is the correctly rounded
quotient for EVERY input; benchmarks/exact_div_check.c tries every fp32 significand against every divisor."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    if shutil.which("gcc") is None:
        pytest.skip("needs gcc")
    exe = str(tmp_path_factory.mktemp("div") / "exact_div_check")
    subprocess.run(["gcc", "-O2", "-mfma", "-fopenmp", "-ffp-contract=off",
                    os.path.join(ROOT, "benchmarks", "exact_div_check.c"), "-o", exe, "-lm"], check=True)
    return exe


@pytest.mark.parametrize("lo,hi", [
    (1, 64),          # every small image side
    (103, 103), (207, 207), (255, 255), (383, 383), (415, 415), (447, 447), (831, 831), (1023, 1023),  # KITTI / Sintel sizes - 1
    (16390, 16400),   # the top of the range the kernels use it for
])
def test_fma_quotient_is_the_ieee_quotient(checker, lo, hi):
    r = subprocess.run([checker, str(lo), str(hi)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.strip() == "bad=0", r.stdout + r.stderr
