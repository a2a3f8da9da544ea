"""bench.py without a GPU: the reference arm (the reference's own compute_using_cpu through oracle/_ref, or the
oracle port) prints exactly ONE JSON line with the keys the contract names, the product arm fails loudly
instead of falling back to the CPU, and the small host helpers (peak lookup, ncu traffic lookup, per-format
table) behave."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT


def run_bench(*args, env=None, timeout=600):
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], capture_output=True, text=True, timeout=timeout,
                          cwd=str(ROOT), env=dict(os.environ, **(env or {})))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    p = run_bench("--impl", "reference", "--gpus", "1", "--steps", "2", "--warmup", "1", "--cpu-sample-rows", "8192")
    assert p.returncode == 0, p.stderr
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, p.stdout                      # the reference's own printf()s must not reach stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "GFLOP/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0
    assert d["dtype"] == "f64"                             # the reference has no fp32
    assert d["config"]["formats"] == ["coo", "csr", "ell", "sell", "cmrs"] and "banded" in d["config"]["workload"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert set(cb["per_format_gflops"]) == {"coo", "csr", "ell", "sell", "cmrs"}
    assert d["e2e"] == {"value": d["value"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_under_a_multi_rank_launch_only_rank_zero_works():
    p = run_bench("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--cpu-sample-rows", "4096",
                  env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert p.returncode == 0 and p.stdout.strip() == "", (p.stdout, p.stderr)


def test_product_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    p = run_bench("--steps", "1", "--warmup", "0", "--no-extras", "--no-e2e", "--no-cpu-baseline", timeout=300)
    assert p.returncode != 0 and p.stdout.strip() == ""     # no JSON line, no CPU fallback
    assert "NO_DEVICE" in p.stderr or "no CUDA" in p.stderr or "B200Error" in p.stderr, p.stderr[-500:]


def test_host_helpers():
    sys.path.insert(0, str(ROOT))
    import bench
    peak, src = bench.measured_peak()
    assert peak > 1000 and ("measured" in src or "fallback" in src)
    assert bench.ncu_traffic("banded", "f32", "cmrs", 2097152) > 1.5e9          # the committed capture
    assert bench.ncu_traffic("laplace-iter", "f64", "sell_fused", 8000000) > 7e8
    assert bench.ncu_traffic("banded", "f32", "cmrs", 12345) is None           # a size that was not captured
    fm = bench.format_table(["a"], {"a": 0.5}, {"a": 2_000_000_000}, 100_000_000, 4000.0)
    assert fm["a"]["gbs"] == 4000.0 and fm["a"]["frac_measured"] == 1.0 and fm["a"]["gflops"] == 400.0
    s = bench.ClockSampler(0).summary()                    # no NVML here: an empty, well-formed summary
    assert set(s) == {"sm_mhz", "sm_max_mhz", "reasons", "samples"} and (s["samples"] == 0 or s["sm_mhz"] > 0)
    assert np.isfinite(bench.NOMINAL_HBM_GBS)
