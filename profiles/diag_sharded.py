"""Where does a sharded step spend its time?  Phase-by-phase host timing (each phase closed by a synchronize) next to
the unsynchronised step.  usage: torchrun --nproc-per-node N profiles/diag_sharded.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from site_analysis_b200 import AssignmentEngine, _native as nat, distributed as sd, synthetic as syn

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
F = 125000
g = syn.argyrodite_geometry(2)
d_frac, d_lat = syn.argyrodite_trajectory(g, F, seed=20260105 + 7919 * rank, device=dev)
first = [d_frac[0].clone(), d_lat[0].clone()]
dist.broadcast(first[0], 0); dist.broadcast(first[1], 0)
frame0, lat0 = first[0].cpu().numpy(), first[1].cpu().numpy()
eng = AssignmentEngine("polyhedral", g.n_atoms, g.mobile_idx, device=local, **syn.polyhedral_engine_kwargs(g, frame0, lat0))
G = dist.group.WORLD


def sync():
    torch.cuda.synchronize()


def phases():
    t = [time.perf_counter()]
    sd._initial_state(eng, frame0, lat0)
    sd._carry_by_anchor(eng, d_frac, G); sync(); t.append(time.perf_counter())
    loc = eng.assign_device(d_frac, d_lat, defer_resolve=True); sync(); t.append(time.perf_counter())
    M = len(eng.fw_atoms)
    note = torch.zeros(6 * M + 1, dtype=torch.float64, device=dev)
    nat.check(eng.L.sab_get_excursion_device(eng.h, note.data_ptr(), torch.cuda.current_stream(dev).cuda_stream))
    parts = torch.stack(sd._all_gather(note, G))
    _ = parts[:, -1].max().cpu(); sync(); t.append(time.perf_counter())
    eng._pending = []
    full = sd.gather_assignments(loc, G, True); sync(); t.append(time.perf_counter())
    return np.diff(t) * 1e3


for _ in range(3):
    phases()
p = np.mean([phases() for _ in range(5)], axis=0)
for mode in ("sync", "async"):
    for _ in range(2):
        sd.assign_sharded(eng, d_frac, d_lat, frame0, lat0, G, equal_shards=True)
    sync(); dist.barrier(); sync()
    t0 = time.perf_counter()
    pend = None
    for _ in range(5):
        if mode == "sync":
            sd.assign_sharded(eng, d_frac, d_lat, frame0, lat0, G, equal_shards=True)
        else:
            full, work = sd.assign_sharded(eng, d_frac, d_lat, frame0, lat0, G, equal_shards=True, async_gather=True)
            if pend is not None:
                pend.wait()
            pend = work
    if pend is not None:
        pend.wait()
    sync()
    dt = (time.perf_counter() - t0) / 5 * 1e3
    if rank == 0:
        print(f"{mode} step: {dt:.3f} ms")
# single-GPU style call on the same engine for comparison
for _ in range(2):
    eng.assign_device(d_frac, d_lat); eng._pending = []
sync(); t0 = time.perf_counter()
for _ in range(5):
    eng.assign_device(d_frac, d_lat); eng._pending = []
sync()
if rank == 0:
    print(f"local assign_device only: {(time.perf_counter() - t0) / 5 * 1e3:.3f} ms; kernel {eng.last_info['assign_ns'] * 1e-6:.3f} ms")
    print("phases (ms, each closed by a synchronize): carry %.3f | assign_device %.3f | guard/flag all-gather %.3f | result all-gather %.3f" % tuple(p))
dist.destroy_process_group()
