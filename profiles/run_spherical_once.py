import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from site_analysis_b200 import AssignmentEngine, synthetic as syn
F = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
g = syn.argyrodite_geometry(2)
d_frac, d_lat = syn.argyrodite_trajectory(g, F, device="cuda")
eng = AssignmentEngine("spherical", g.n_atoms, g.mobile_idx, centres=g.site_centres, rcut=1.5)
eng.assign_device(d_frac, d_lat); torch.cuda.synchronize()
print("ok", eng.last_info)
