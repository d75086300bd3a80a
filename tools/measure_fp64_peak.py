#!/usr/bin/env python3
"""The FP64 roofline denominator, measured: `sadmc_measure_fp64_peak` (independent DFMA chains, csrc/engine.cu) with the SM clock
sampled while it runs -> profiles/r02_fp64_peak.json.  MEASURED_PEAKS.json carries no FP64 figure; this file is the record of
where bench.py's `roofline.peak` comes from and of the clock it was taken at."""
import ctypes as C
import json
import os
import subprocess
import sys
import threading

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sad_monte_carlo_b200 import load_library  # noqa: E402

lib = load_library()
samples, stop = [], threading.Event()


def sample():
    while not stop.is_set():
        try:
            o = subprocess.run(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active",
                                "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout.strip()
            samples.append([x.strip() for x in o.split(",")])
        except Exception:
            pass
        stop.wait(0.1)


t = threading.Thread(target=sample, daemon=True)
t.start()
vals = []
for _ in range(5):
    v = C.c_double()
    assert lib.sadmc_measure_fp64_peak(0, 20, C.byref(v)) == 0
    vals.append(v.value)
stop.set()
t.join(timeout=2)
clk = sorted(float(s[0]) for s in samples if s and s[0].replace(".", "").isdigit())
out = {"fp64_tflops_best_of_5x20": max(vals), "all": vals, "sm_mhz_median_under_load": clk[len(clk) // 2] if clk else None,
       "sm_max_mhz": float(samples[0][1]) if samples else None, "clock_samples": len(clk),
       "nominal_at_max_clock": "148 SMs x 64 DFMA/clk x 2 flop x sm_max_mhz",
       "nominal_tflops": 148 * 64 * 2 * float(samples[0][1]) * 1e6 / 1e12 if samples else None,
       "how": "8 independent DFMA chains per thread, 256 threads x 8 blocks per SM, 4096 x 64 DFMA per thread, CUDA events, best of 20 launches"}
json.dump(out, open(os.path.join(ROOT, "profiles", "r02_fp64_peak.json"), "w"), indent=1)
print(json.dumps(out))
