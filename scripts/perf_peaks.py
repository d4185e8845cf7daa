#!/usr/bin/env python
"""scripts/perf_peaks.py -- the roofline denominators bench.py measures live: FP32 FMA TFLOP/s and shared-memory / L1 / L2
read GB/s (skr_measure_fp32_peak, skr_measure_bandwidth)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import skele_raytracer_b200 as S  # noqa: E402

r = S.Renderer()
print("fp32 TFLOP/s", r.measure_fp32_peak(4096), "lds GB/s", r.measure_bandwidth(0), "l1 GB/s", r.measure_bandwidth(1), "l2 GB/s", r.measure_bandwidth(2), flush=True)
