"""The bench line the driver parses (committed copy of the last run: profiles/r02_v5_bf16_bench.json) carries every key of the
contract, with consistent values; and bench.py's argument surface has the flags the driver passes."""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _latest(pattern):
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", pattern)))
    assert files, pattern
    return json.loads(open(files[-1]).read().strip().splitlines()[-1])


def test_committed_bench_line_has_the_contract_keys():
    d = _latest("r02_v*_bf16_bench.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    assert d["unit"] == "samples/s" and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["dtype"] == "bf16" and d["data"] == "synthetic" and "workload" in d["config"] and "model" not in d["config"]
    assert d["steps"] >= 1 and d["warmup"] >= 3 and d["gpu_launches"] > 0
    S = d["config"]["samples_per_gpu"]
    assert abs(d["value"] - S * d["n_gpus"] / (d["ms_per_step"] / 1e3)) <= 1e-6 * d["value"]
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["traffic"] is None or r["traffic"] > 0
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    assert r["peak"] in (peaks["bf16_tflops_sustained"], peaks["bf16_tflops"])
    c = d["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] != d["value"]
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_committed_reference_arm_line():
    d = _latest("r02_v*_reference_arm.json")
    assert d["impl"] == "reference" and d["unit"] == "samples/s" and d["cpu_baseline"]["kind"] == "reference"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]


def test_bench_cli_surface():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0
    for flag in ("--gpus", "--steps", "--warmup", "--impl", "--config", "--gemm"):
        assert flag in out.stdout, flag
