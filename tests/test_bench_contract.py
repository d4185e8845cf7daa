"""bench.py's output contract, checked without a GPU:
  * the committed bench lines under profiles/ (measured on B200s) carry every key the driver and the judge read;
  * `bench.py --impl reference` (the reference's own CPU code when oracle/_ref is built, else the C port) runs here and
    prints exactly ONE JSON line on stdout with the reference-arm keys.
"""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

PROFILES = os.path.join(ROOT, "profiles")
LINE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
             "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"}


def load(name):
    lines = [ln for ln in open(os.path.join(PROFILES, name)).read().splitlines() if ln.strip()]
    assert len(lines) == 1, f"{name}: one JSON line expected, got {len(lines)}"
    return json.loads(lines[0])


@pytest.mark.parametrize("name,n", [("r01_bench_n1.json", 1), ("r01_bench_n2.json", 2), ("r01_bench_n4.json", 4), ("r01_bench_n8.json", 8)])
def test_committed_bench_lines_keep_the_contract(name, n):
    d = load(name)
    assert LINE_KEYS <= set(d), LINE_KEYS - set(d)
    assert d["metric"] == "Mrays/s" and d["unit"] == "Mrays/s" and d["higher_is_better"] is True and d["n_gpus"] == n
    assert d["warmup"] >= 3 and d["steps"] >= 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["scaling"] == "strong" and d["vs_baseline"] is None and d["dtype"] == "f32" and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"] and "l2" in d["config"]
    # value is whole-job throughput: rays of the frame / step time
    assert d["value"] == pytest.approx(d["config"]["rays_per_frame"] / (d["ms_per_step"] * 1e-3) / 1e6, rel=1e-6)
    c = d["clocks"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(c) and not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    e = d["e2e"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(e)
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] == d["config"]["width"] * d["config"]["height"] * 3
    assert 0 < e["value"] < d["value"]                      # end to end can only be slower than the resident-data figure
    assert d["gpu_launches"] >= d["steps"]
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic", "executed"} <= set(r)
    assert r["frac"] == pytest.approx(r["achieved"] / r["peak"]) and 0 < r["executed"]["frac"] <= r["frac"] < 1
    assert r["executed"]["sphere_tests_executed"] <= r["executed"]["sphere_tests_algorithmic"]
    if n == 1:
        b = d["cpu_baseline"]
        assert {"value", "unit", "cores", "kind", "sample"} <= set(b) and b["kind"] in ("reference", "port") and b["cores"] >= 1
        assert isinstance(r["traffic"], int) and r["traffic"] > 0
    else:
        assert d["cpu_baseline"] is None                    # timed on rank 0 at N = 1 only


def test_strong_scaling_is_monotonic_in_the_committed_lines():
    ms = [load(f"r01_bench_n{n}.json")["ms_per_step"] for n in (1, 2, 4, 8)]
    assert ms[0] > ms[1] > ms[2] > ms[3]
    assert ms[0] / ms[3] > 4.0                              # 8 GPUs on a 1 ms frame


def test_reference_arm_runs_on_the_host_and_prints_one_json_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"], capture_output=True,
                       text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, p.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "Mrays/s" and d["unit"] == "Mrays/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["gpu_launches"] == 0 and d["config"]["workload"] == load("r01_bench_n1.json")["config"]["workload"]
    b = d["cpu_baseline"]
    assert b["kind"] in ("reference", "port") and b["cores"] >= 1 and b["value"] == d["value"] and b["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
