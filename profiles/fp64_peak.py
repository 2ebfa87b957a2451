"""Measured FP64 issue peaks of the device without FMA (the denominator for the float64 kernels whose decisions must be
bit-identical to numba code that never fuses multiply-add).  Writes gpurun_out/fp64_peak.json.
usage (GPU box): python profiles/fp64_peak.py"""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from site_analysis_b200 import _native as nat

t = np.zeros(4)
nat.check(nat.lib().sab_measure_fp64_peak(0, nat.ptr(t)))
out = {"dadd_tflops": t[0], "dmul_tflops": t[1], "dmul_dadd_pairs_tflops": t[2], "dfma_tflops_2flop": t[3],
       "how": "148 x 8 CTAs x 256 threads, 8 independent accumulators per thread, 20000 iterations, best of 3 (CUDA events)"}
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/fp64_peak.json", "w"), indent=1)
print(json.dumps(out))
