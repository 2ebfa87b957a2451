"""cProfile of the user-facing call Trajectory.trajectory_from_arrays on the bench workload (pinned host buffers).
usage (GPU box): python profiles/prof_api.py [frames]"""
import cProfile
import os
import pstats
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

F = int(sys.argv[1]) if len(sys.argv) > 1 else 125000
wl = bench.Workload("polyhedral")
d_frac, d_lat = wl.trajectory(F, wl.seed, "cuda")
h_frac = torch.empty(d_frac.shape, dtype=torch.float64, pin_memory=True); h_frac.copy_(d_frac)
h_lat = torch.empty(d_lat.shape, dtype=torch.float64, pin_memory=True); h_lat.copy_(d_lat)
nf, nl = h_frac.numpy(), h_lat.numpy()
del d_frac, d_lat
traj = wl.api_trajectory(nf[0], nl[0])
traj.trajectory_from_arrays(nf, nl)
torch.cuda.synchronize()
for _ in range(2):
    t0 = time.perf_counter(); traj.trajectory_from_arrays(nf, nl); torch.cuda.synchronize()
    print(f"call: {1e3 * (time.perf_counter() - t0):.1f} ms")
pr = cProfile.Profile()
pr.enable(); traj.trajectory_from_arrays(nf, nl); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
