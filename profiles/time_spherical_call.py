import sys, time, os
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from site_analysis_b200 import AssignmentEngine, synthetic as syn
g = syn.argyrodite_geometry(2)
d_frac, d_lat = syn.argyrodite_trajectory(g, 20000, device="cuda")
eng = AssignmentEngine("spherical", g.n_atoms, g.mobile_idx, centres=g.site_centres, rcut=1.5)
for i in range(3):
    torch.cuda.synchronize(); t0=time.perf_counter()
    eng.assign_device(d_frac, d_lat); eng._pending=[]
    torch.cuda.synchronize(); print('call', time.perf_counter()-t0, eng.last_info)
import cProfile, pstats
pr=cProfile.Profile(); pr.enable(); eng.assign_device(d_frac, d_lat); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(14)
