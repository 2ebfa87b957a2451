"""Timing of the stand-alone minimum-image distance operator (sab_mic_distances, csrc/k_distances.cuh).
usage (GPU box): python profiles/bench_distances.py > gpurun_out/distances.json
3328 x 3328 pairs (4x4x4 argyrodite supercell against itself, the alignment / mapping matrix of SURVEY 8(f) rank 3) and
8448 x 1536 (all sites x all mobile ions of BASELINE config 3): device time by CUDA events, FP64 rate counted as SURVEY
8(d) does (369 flop per pair: no FMA, each add / mul / compare = 1), bytes written, and the CPU oracle on one core."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy as np
import torch

from site_analysis_b200.distances import all_mic_distances_device

dev = torch.device("cuda", 0)
lat = torch.tensor([[40.62, 1e-4, 0.0], [2e-4, 40.7, -1e-4], [0.3, 0.0, 40.5]], dtype=torch.float64, device=dev)
out = {}
for n, m in ((3328, 3328), (8448, 1536)):
    g = torch.Generator(device=dev).manual_seed(n + m)
    c1 = torch.rand((n, 3), dtype=torch.float64, device=dev, generator=g)
    c2 = torch.rand((m, 3), dtype=torch.float64, device=dev, generator=g)
    for _ in range(3):
        d = all_mic_distances_device(c1, c2, lat)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        d = all_mic_distances_device(c1, c2, lat)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    rec = dict(pairs=n * m, kernel_ms=ms, pairs_per_s=n * m / (ms * 1e-3), fp64_tflops_nonfma=369 * n * m / (ms * 1e-3) / 1e12,
               write_gbs=8 * n * m / (ms * 1e-3) / 1e9)
    import oracle
    k = min(n, 256)
    t0 = time.perf_counter()
    want = oracle.all_mic_distances(c1[:k].cpu().numpy(), c2.cpu().numpy(), lat.cpu().numpy())
    rec["cpu_oracle_pairs_per_s_1core"] = k * m / (time.perf_counter() - t0)
    rec["bit_identical_to_oracle_on_sample"] = bool(np.array_equal(d[:k].cpu().numpy().view(np.int64), want.view(np.int64)))
    out[f"{n}x{m}"] = rec
print(json.dumps(out))
