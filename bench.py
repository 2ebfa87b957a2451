#!/usr/bin/env python
"""Headline benchmark: polyhedral ion-frame assignments per second (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (one rank per GPU)
    python bench.py --impl reference --steps K --warmup W    # CPU arm: the oracle port, all host cores

Workload (config.workload): BASELINE.json config 5 per GPU -- the argyrodite Li6PS5Cl 2x2x2
geometry of config 2 (N = 416 atoms, A = 192 Li, S = 1056 S4 tetrahedra) with synthetic
dynamics, 125,000 frames per GPU (1,000,000 frames over 8 GPUs; weak scaling).  One "step" is
one pass of the hot path over the rank's 125k-frame shard; successive steps continue the same
periodic trajectory, so the PBC wrap state carries on exactly as in a longer run.

value   device-resident: frames already in HBM, timed with CUDA events, max over ranks.
e2e     the C-ABI host-buffer call (sab_assign_host): pinned host frames in, host int32 out,
        H2D/D2H inside the timed region.
The shard (1.25 GB) is ~10x the 126 MB L2, so no explicit L2 flush is needed between steps.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "ion-frame assignments/sec (polyhedral)"
UNIT = "ion-frames/s"
FRAMES_PER_GPU = 125_000


def algorithmic_bytes_per_frame(n_mobile: int, n_framework_referenced: int) -> int:
    """mobile + referenced framework coordinates and the lattice in,
    one int32 site id per mobile atom out."""
    return 24 * (n_mobile + n_framework_referenced) + 72 + 4 * n_mobile


def measured_hbm_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi SM clock / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self) -> dict:
        self.stop_flag.set()
        self.join(timeout=6)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(s) > 2 + i and s[2 + i] == "Active" for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": int(self.samples[0][1]) if self.samples[0][1].isdigit() else None, "reasons": reasons}


def oracle_for(geom, frame0, lattice0):
    """The CPU checker on the same site tables (test infrastructure; never on the product path)."""
    sys.path.insert(0, str(ROOT / "oracle"))
    import oracle
    from site_analysis_b200 import _tables as tab, synthetic as syn
    kw = syn.polyhedral_engine_kwargs(geom, frame0, lattice0)
    S = len(kw["slots"])
    topo = []
    for t, (sh, raw) in zip(kw["slots"], kw["primed"]):
        topo.append(tab.hull_simplices(tab.unwrapped_vertices(raw, sh)))
    make = lambda: oracle.SiteOracle(                                    # noqa: E731
        "polyhedral", geom.mobile_idx, vert_idx=np.array(kw["slots"], dtype=np.int32), n_vert=np.full(S, 4, np.int32),
        has_ref_center=np.ones(S, np.uint8), ref_center=np.array(kw["reference_centres"]), primed=np.ones(S, np.uint8),
        shift0=np.array([p[0] for p in kw["primed"]]), cached_raw0=np.array([p[1] for p in kw["primed"]]), topology=topo)
    return make


def time_oracle_all_cores(make, frac, lat, frames_per_thread):
    """All host cores: one independent oracle replica per thread on a contiguous frame slice
    (SURVEY.md section 8(d) CPU baseline (ii)); throughput = ion-frames / wall time of the slowest."""
    cores = os.cpu_count() or 1
    n = min(cores, max(1, frac.shape[0] // frames_per_thread))
    oracles = [make() for _ in range(n)]
    slices = [(i * frames_per_thread, (i + 1) * frames_per_thread) for i in range(n)]
    for o, (a, b) in zip(oracles, slices):                  # one warm-up frame each (caches, topology)
        o.run(frac[a:a + 1], lat[a:a + 1])
    threads = [threading.Thread(target=o.run_timed, args=(frac[a:b], lat[a:b])) for o, (a, b) in zip(oracles, slices)]
    t0 = time.perf_counter()
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    dt = time.perf_counter() - t0
    return n * frames_per_thread * oracles[0].A / dt, n, dt


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from site_analysis_b200 import synthetic as syn
    geom = syn.argyrodite_geometry(2)
    cores = os.cpu_count() or 1
    per_thread = args.ref_frames_per_thread
    frac, lat = syn.argyrodite_trajectory(geom, cores * per_thread, seed=20260105, device="cpu")
    frac, lat = frac.numpy(), lat.numpy()
    make = oracle_for(geom, frac[0], lat[0])
    for _ in range(args.warmup):
        time_oracle_all_cores(make, frac, lat, per_thread)
    vals, t_tot = [], 0.0
    for _ in range(args.steps):
        v, n, dt = time_oracle_all_cores(make, frac, lat, per_thread)
        vals.append(v); t_tot += dt
    value = float(np.mean(vals))
    sample = f"{n} threads x {per_thread} frames of the synthetic argyrodite trajectory per step"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(geom, per_thread * n),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": n, "kind": "port", "sample": sample,
                         "note": "plain-C oracle port of the reference's sequential algorithm (oracle/site_oracle.c); the "
                                 "reference itself is Python+numba and cannot travel to the GPU box"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(geom, frames_per_gpu):
    return {"workload": "synthetic argyrodite Li6PS5Cl 2x2x2 polyhedral (BASELINE config 5 shard: config-2 geometry, "
                        "synthetic dynamics)", "frames_per_gpu": int(frames_per_gpu), "atoms_per_frame": geom.n_atoms,
            "mobile_atoms": geom.n_mobile, "sites": len(geom.site_centres), "vertices_per_site": 4,
            "l2": "inputs larger than L2 (1.25 GB shard vs 126 MB)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=FRAMES_PER_GPU)
    ap.add_argument("--ref-frames-per-thread", type=int, default=10000)
    ap.add_argument("--parity-frames", type=int, default=400)
    ap.add_argument("--cpu-baseline-frames", type=int, default=0,
                    help="frames per host thread of the cpu_baseline sample (0 = the whole shard split over all cores)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    from site_analysis_b200 import AssignmentEngine, synthetic as syn
    # ONE JSON line on stdout: native libraries (NCCL's version banner ...) write to file descriptor 1 directly, so it is
    # pointed at stderr for the run and the line goes to the saved descriptor at the end
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    F = args.frames
    geom = syn.argyrodite_geometry(2)
    A, N = geom.n_mobile, geom.n_atoms
    d_frac, d_lat = syn.argyrodite_trajectory(geom, F, seed=20260105 + 7919 * rank, device=dev)
    frame0, lat0 = d_frac[0].cpu().numpy(), d_lat[0].cpu().numpy()
    eng = AssignmentEngine("polyhedral", N, geom.mobile_idx, device=local,
                           **syn.polyhedral_engine_kwargs(geom, frame0, lat0))
    if world > 1:
        # every rank primes its site tables from the trajectory's first frame (rank 0's frame 0)
        first = [torch.empty_like(d_frac[0]), torch.empty_like(d_lat[0])]
        if rank == 0:
            first = [d_frac[0].clone(), d_lat[0].clone()]
        dist.broadcast(first[0], 0); dist.broadcast(first[1], 0)
        frame0, lat0 = first[0].cpu().numpy(), first[1].cpu().numpy()
        eng = AssignmentEngine("polyhedral", N, geom.mobile_idx, device=local,
                               **syn.polyhedral_engine_kwargs(geom, frame0, lat0))
        from site_analysis_b200 import distributed as sd

    in_flight = []                                         # (full result, NCCL work) of the previous step
    peer_bufs, exchange = None, "single GPU"
    if world > 1:
        try:     # result rows stored by the kernel itself into every rank's full array (NVLink peer memory)
            if os.environ.get("SAB_NO_PEER_STORE") == "1":      # A/B switch: NCCL all-gather instead
                raise RuntimeError("disabled by SAB_NO_PEER_STORE=1")
            peer_bufs = sd.PeerResultBuffers(F, A, dist.group.WORLD, dev)
            exchange = "fused: the stream kernel stores its result rows into every rank's full array over NVLink peer memory"
        except Exception as exc:
            exchange = f"NCCL all-gather overlapping the next step's kernels (symmetric memory unavailable: {type(exc).__name__})"

    def drain():
        while in_flight:
            full, work = in_flight.pop()
            work.wait()                                    # orders the current stream after that step's all-gather
        return None

    def step():
        if world > 1:
            # shard carry (tiny all-gather) + local kernels + guard/flag all-gather + the one result all-gather.  A long
            # trajectory streams through in batches: the result all-gather of batch k runs while batch k + 1 computes
            # (it is waited for one step later, and all of them before the timed region closes).
            full, work = sd.assign_sharded(eng, d_frac, d_lat, frame0, lat0, dist.group.WORLD, equal_shards=True,
                                           async_gather=True, peer_buffers=peer_bufs)
            drain()
            if work is not None:
                in_flight.append((full, work))
            return full[rank * F:(rank + 1) * F]
        res = eng.assign_device(d_frac, d_lat)
        eng._pending = []                                  # array path: no Python-object history to maintain
        return res

    for _ in range(max(args.warmup, 0)):
        res = step()
    drain()
    launches, assign_ns, wrap_ns = 0, 0, 0
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        res = step()
        launches += eng.last_info["launches"] + (2 if world > 1 else 0)
        assign_ns += eng.last_info["assign_ns"]; wrap_ns += eng.last_info["wrap_ns"]
    drain()                                                # every result all-gather has finished inside the timed region
    ev1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    clocks = sampler.summary() if sampler else None
    value = world * F * A * args.steps / (ms_total * 1e-3)

    # ---- parity spot check against the oracle (outside every timed region) -------------------
    parity = None
    if rank == 0 and args.parity_frames > 0:
        npar = min(args.parity_frames, F)
        eng_p = AssignmentEngine("polyhedral", N, geom.mobile_idx, device=local,
                                 **syn.polyhedral_engine_kwargs(geom, frame0, lat0))
        got = eng_p.assign_device(d_frac[:npar].contiguous(), d_lat[:npar].contiguous()).cpu().numpy()
        want = oracle_for(geom, frame0, lat0)().run(d_frac[:npar].cpu().numpy(), d_lat[:npar].cpu().numpy(), positions=True)
        parity = {"frames": int(npar), "mismatches": int((got != want).sum())}

    # ---- end to end through the C-ABI host call (rank-local shard; pinned host buffers) --------
    h_frac = torch.empty(d_frac.shape, dtype=torch.float64, pin_memory=True); h_frac.copy_(d_frac)
    h_lat = torch.empty(d_lat.shape, dtype=torch.float64, pin_memory=True); h_lat.copy_(d_lat)
    h_out = torch.empty((F, A), dtype=torch.int32, pin_memory=True)
    eng_h = AssignmentEngine("polyhedral", N, geom.mobile_idx, device=local,
                             **syn.polyhedral_engine_kwargs(geom, frame0, lat0))
    nf, nl, no = h_frac.numpy(), h_lat.numpy(), h_out.numpy()
    for _ in range(min(args.warmup, 2)):
        eng_h.assign_host_pipelined(nf, nl, out=no); eng_h._pending = []
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        eng_h.assign_host_pipelined(nf, nl, out=no); eng_h._pending = []
    torch.cuda.synchronize()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = world * F * A * args.steps / float(e2e_s.item())
    same = bool(np.array_equal(no, res.cpu().numpy())) if rank == 0 else None

    if rank == 0:
        peak, peak_src = measured_hbm_peak()
        n_fw = len(eng.fw_atoms)
        bytes_per_launch = algorithmic_bytes_per_frame(A, n_fw) * F
        assign_s = assign_ns * 1e-9 / args.steps
        achieved = bytes_per_launch / assign_s / 1e9
        cores = os.cpu_count() or 1
        cb_frames = args.cpu_baseline_frames or max(1, F // cores)
        nb = min(F, cores * cb_frames)
        cpu_baseline = None
        if world == 1:                                    # reported at N = 1 only
            h_f, h_l = d_frac[:nb].cpu().numpy(), d_lat[:nb].cpu().numpy()
            runs = [time_oracle_all_cores(oracle_for(geom, frame0, lat0), h_f, h_l, cb_frames) for _ in range(3)]
            cpu_v, cpu_n, cpu_dt = float(np.mean([r[0] for r in runs])), runs[0][1], float(np.sum([r[2] for r in runs]))
            cpu_baseline = {"value": cpu_v, "unit": UNIT, "cores": cpu_n, "kind": "port",
                            "sample": f"3 passes of {cpu_n} threads x {cb_frames} frames of this workload "
                                      f"({cpu_dt:.2f} s wall, {cpu_dt * cpu_n:.0f} s of CPU work)"}
        cfg = workload_config(geom, F)
        if world > 1:
            cfg["multi_gpu"] = ("contiguous frame shards, one per rank; per step: halo all-gather, kernels, guard/flag all-gather "
                                "(completion barrier), full int32[F_total, A] result on every rank; result exchange = " + exchange)
        traffic = None
        tp = ROOT / "profiles" / "r1_traffic.json"
        if tp.exists():                                   # one `ncu --set full` capture of this kernel, scaled per frame
            tj = json.loads(tp.read_text())
            traffic = (tj["dram_bytes_read"] + tj["dram_bytes_write"]) / tj["frames"] * F
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": cfg,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(F * N * 24 + F * 72),
                    "d2h_bytes_per_step": int(F * A * 4), "api": "sab_assign_host (C ABI, pinned host buffers)",
                    "equals_device_result": same},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "k_polyhedral_assign_stream", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "bytes_per_launch": int(bytes_per_launch), "kernel_ms": assign_s * 1e3,
                         "wrap_scan_ms": wrap_ns * 1e-6 / args.steps},
            "cpu_baseline": cpu_baseline,
            "parity": parity,
        }
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
