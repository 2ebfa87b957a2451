#!/usr/bin/env python
"""Headline benchmark: ion-frame assignments per second (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--workload NAME]      # our CUDA path (one rank per GPU)
    python bench.py --impl reference --steps K --warmup W [--workload NAME]   # CPU arm: the reference itself, all host cores

Workloads (config.workload), all synthetic (site_analysis_b200/synthetic.py):
  polyhedral     (default, the metric's configuration) BASELINE config 5 per GPU: argyrodite Li6PS5Cl 2x2x2 geometry of
                 config 2 (N = 416 atoms, A = 192 Li, S = 1056 S4 tetrahedra), 125,000 frames per GPU (1 M frames over 8 GPUs)
  voronoi_4x4x4  BASELINE config 3: 4x4x4 argyrodite, N = 3328, A = 1536, S = 8448 Voronoi centres, 100,000 frames
  dynvor_llzo    BASELINE config 4: LLZO garnet 2x2x2, N = 1536, A = 448, S = 576 dynamic-Voronoi sites in two centre groups,
                 200,000 frames
  spherical      argyrodite 2x2x2, 1056 spheres of r = 1.5 A (98 % of the atom-frames sit in several spheres: priority resolver)

One "step" is one pass of the hot path over the rank's shard; successive steps continue the same periodic trajectory, so the
PBC wrap state carries on exactly as in a longer run.

value               device-resident: frames already in HBM, timed with CUDA events, max over ranks.
value_with_history  the same step plus the history fold (transition Counters, recent sites) that the reference keeps per
                    frame (site_collection.py:279-310) -- csrc/k_fold.cuh.
e2e                 the C-ABI host-buffer call (sab_assign_host): pinned host frames in, host int32 out, H2D/D2H inside.
e2e_api             the user-facing call Trajectory.trajectory_from_arrays on the same pinned host buffers.
cpu_baseline        the unmodified reference (oracle/_ref, Python + numba) on all host cores, on a bounded sample (kind
                    "reference"); the plain-C oracle port beside it.  --impl reference prints the same as its own line.
Shards are several times the 126 MB L2, so no explicit L2 flush is needed between steps.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

UNIT = "ion-frames/s"


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


# ------------------------------------------------------------------------------------------------ workloads
class Workload:
    """One benchmark configuration: geometry, site tables for our engine, for the oracle port and for the reference."""

    def __init__(self, name: str):
        from site_analysis_b200 import synthetic as syn
        self.name = name
        self.rcut = None
        if name == "polyhedral":
            self.geom, self.kind, self.frames, self.seed = syn.argyrodite_geometry(2), "polyhedral", 125_000, 20260105
            self.metric = "ion-frame assignments/sec (polyhedral)"
            self.slots = [list(map(int, t)) for t in self.geom.tetrahedra]
            self.kernel, self.ref_frames, self.port_frames = "k_polyhedral_assign_stream5", 300, 8000
            self.label = ("synthetic argyrodite Li6PS5Cl 2x2x2 polyhedral (BASELINE config 5 shard: config-2 geometry, "
                          "synthetic dynamics)")
        elif name == "voronoi_4x4x4":
            self.geom, self.kind, self.frames, self.seed = syn.argyrodite_geometry(4), "voronoi", 100_000, 20260103
            self.metric = "ion-frame assignments/sec (voronoi)"
            self.slots = None
            self.kernel, self.ref_frames, self.port_frames = "k_nearest_assign_cells", 3, 6
            self.label = "synthetic argyrodite 4x4x4, Voronoi sites with PBC (BASELINE config 3)"
        elif name == "dynvor_llzo":
            self.geom, self.kind, self.frames, self.seed = syn.garnet_geometry(2), "dynamic_voronoi", 200_000, 20260104
            self.metric = "ion-frame assignments/sec (dynamic voronoi)"
            self.slots = [list(map(int, s)) for s in self.geom.reference_slots]
            self.kernel, self.ref_frames, self.port_frames = "k_dynvor_assign_cells", 60, 400
            self.label = "synthetic LLZO garnet 2x2x2, dynamic Voronoi sites in two centre groups (BASELINE config 4)"
        elif name == "spherical":
            self.geom, self.kind, self.frames, self.seed = syn.argyrodite_geometry(2), "spherical", 20_000, 20260105
            self.metric = "ion-frame assignments/sec (spherical)"
            self.slots, self.rcut = None, 1.5
            self.kernel, self.ref_frames, self.port_frames = "k_spherical_assign_cells", 1000, 8000
            self.label = "synthetic argyrodite 2x2x2, 1056 spherical sites r = 1.5 A (overlapping: priority resolver)"
        else:
            raise SystemExit(f"unknown workload {name}")
        g = self.geom
        self.A, self.N, self.S = g.n_mobile, g.n_atoms, len(g.site_centres)
        self.n_fw = len({a for s in self.slots for a in s}) if self.slots else 0
        # mobile + referenced framework coordinates and the lattice in, one int32 per mobile atom out
        self.bytes_per_frame = 24 * (self.A + self.n_fw) + 72 + 4 * self.A

    def trajectory(self, n_frames, seed, device):
        from site_analysis_b200 import synthetic as syn
        return syn.synthetic_trajectory(self.geom, n_frames, seed=seed, device=device)

    def engine_kwargs(self, frame0, lat0):
        from site_analysis_b200 import synthetic as syn
        g = self.geom
        if self.kind == "polyhedral":
            return syn.polyhedral_engine_kwargs(g, frame0, lat0)
        if self.kind == "voronoi":
            return dict(centres=g.site_centres)
        if self.kind == "spherical":
            return dict(centres=g.site_centres, rcut=self.rcut)
        return dict(slots=self.slots, reference_centres=list(g.site_centres))

    def engine(self, frame0, lat0, device):
        from site_analysis_b200 import AssignmentEngine
        return AssignmentEngine(self.kind, self.N, self.geom.mobile_idx, device=device, **self.engine_kwargs(frame0, lat0))

    def oracle(self, frame0, lat0):
        """The CPU checker on the same site tables (test infrastructure; never on the product path)."""
        sys.path.insert(0, str(ROOT / "oracle"))
        import oracle
        from site_analysis_b200 import _tables as tab
        g, S = self.geom, self.S
        if self.kind == "polyhedral":
            kw = self.engine_kwargs(frame0, lat0)
            topo = [tab.hull_simplices(tab.unwrapped_vertices(raw, sh)) for sh, raw in kw["primed"]]
            return oracle.SiteOracle(
                "polyhedral", g.mobile_idx, vert_idx=np.array(kw["slots"], dtype=np.int32), n_vert=np.full(S, 4, np.int32),
                has_ref_center=np.ones(S, np.uint8), ref_center=np.array(kw["reference_centres"]), primed=np.ones(S, np.uint8),
                shift0=np.array([p[0] for p in kw["primed"]]), cached_raw0=np.array([p[1] for p in kw["primed"]]), topology=topo)
        if self.kind == "voronoi":
            return oracle.SiteOracle("voronoi", g.mobile_idx, centres=g.site_centres)
        if self.kind == "spherical":
            return oracle.SiteOracle("spherical", g.mobile_idx, centres=g.site_centres, rcut=self.rcut)
        width = max(map(len, self.slots))
        ref = np.zeros((S, width), dtype=np.int32)
        for i, s in enumerate(self.slots):
            ref[i, :len(s)] = s
        return oracle.SiteOracle("dynamic_voronoi", g.mobile_idx, ref_idx=ref, n_ref=np.array([len(s) for s in self.slots], np.int32),
                                 has_ref_center=np.ones(S, np.uint8), ref_center=g.site_centres)

    def config(self, frames_per_gpu):
        return {"workload": self.label, "name": self.name, "frames_per_gpu": int(frames_per_gpu), "atoms_per_frame": self.N,
                "mobile_atoms": self.A, "sites": self.S, "site_kind": self.kind,
                "l2": f"inputs larger than L2 ({frames_per_gpu * self.N * 24 / 1e9:.2f} GB shard vs 126 MB)"}

    def api_trajectory(self, frame0, lat0):
        """The reference-shaped objects a user would build (Trajectory of sites and atoms), primed from the first frame."""
        import site_analysis_b200 as sab
        g = self.geom
        sab.Site.reset_index()
        if self.kind == "polyhedral":
            sites = [sab.PolyhedralSite([int(v) for v in t], reference_center=c) for t, c in zip(self.slots, g.site_centres)]
            for s in sites:
                s.assign_vertex_coords(frame0, lat0)
        elif self.kind == "voronoi":
            sites = [sab.VoronoiSite(c) for c in g.site_centres]
        elif self.kind == "spherical":
            sites = [sab.SphericalSite(c, self.rcut) for c in g.site_centres]
        else:
            sites = [sab.DynamicVoronoiSite(s, reference_center=c) for s, c in zip(self.slots, g.site_centres)]
        return sab.Trajectory(sites=sites, atoms=[sab.Atom(index=int(i)) for i in g.mobile_idx])


def oracle_for(geom, frame0, lattice0):
    """Factory of the CPU checker for the polyhedral argyrodite workload on geometry ``geom`` (used by tests/ and profiles/)."""
    wl = Workload("polyhedral")
    wl.geom, wl.slots = geom, [list(map(int, t)) for t in geom.tetrahedra]
    wl.A, wl.N, wl.S = geom.n_mobile, geom.n_atoms, len(geom.site_centres)
    return lambda: wl.oracle(frame0, lattice0)


# ------------------------------------------------------------------------------------------------ CPU arms
def port_all_cores(wl: Workload, frac, lat, frames_per_thread, passes=1):
    """The plain-C oracle port: one independent replica per host thread on a contiguous frame slice (SURVEY.md section 8(d)
    CPU baseline (ii)); throughput = ion-frames / wall time of the slowest."""
    cores = os.cpu_count() or 1
    n = min(cores, max(1, frac.shape[0] // frames_per_thread))
    vals, t_tot = [], 0.0
    for _ in range(passes):
        oracles = [wl.oracle(frac[0], lat[0]) for _ in range(n)]
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
        vals.append(n * frames_per_thread * wl.A / dt); t_tot += dt
    return float(np.mean(vals)), n, t_tot


def reference_available():
    ref = ROOT / "oracle" / "_ref" / "site_analysis"
    if not ref.is_dir():
        return "oracle/_ref is missing (run __graft_entry__.build() where /root/reference is mounted)"
    try:
        import numba  # noqa: F401
    except Exception as exc:                                    # the reference would take its slow non-numba path
        return f"numba is not importable on this host ({type(exc).__name__})"
    return None


def reference_all_cores(wl: Workload, frac, lat, frames_per_proc):
    """The UNMODIFIED reference (oracle/_ref, Python + numba): one process per host core, NUMBA_NUM_THREADS=1, each on an
    equal contiguous frame slice after a one-frame warm-up (tests/benchmark_all_site_types.py:92-103; SURVEY.md section 8(d)
    CPU baseline (ii)).  Throughput = total ion-frames / wall time of the slowest worker's timed loop."""
    cores = os.cpu_count() or 1
    per = frames_per_proc + 1                                   # + the warm-up frame
    n = min(cores, max(1, frac.shape[0] // per))
    base = Path("/dev/shm") if Path("/dev/shm").is_dir() else Path(tempfile.gettempdir())
    with tempfile.TemporaryDirectory(dir=base) as tmp:
        tmp = Path(tmp)
        np.save(tmp / "frac.npy", np.ascontiguousarray(frac[:n * per])); np.save(tmp / "lat.npy", np.ascontiguousarray(lat[:n * per]))
        tables = dict(species=np.array(wl.geom.species), centres=wl.geom.site_centres, mobile=wl.geom.mobile_idx)
        if wl.slots is not None:
            width = max(map(len, wl.slots))
            sl = np.zeros((wl.S, width), dtype=np.int64)
            for i, s in enumerate(wl.slots):
                sl[i, :len(s)] = s
            tables.update(slots=sl, n_ref=np.array([len(s) for s in wl.slots]))
        if wl.rcut is not None:
            tables["rcut"] = np.array(wl.rcut)
        np.savez(tmp / "tables.npz", **tables)
        env = dict(os.environ, NUMBA_NUM_THREADS="1", OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1", MKL_NUM_THREADS="1")
        procs = [subprocess.Popen([sys.executable, str(ROOT / "oracle" / "ref_worker.py"), wl.kind, str(tmp / "tables.npz"),
                                   str(tmp / "frac.npy"), str(tmp / "lat.npy"), str(i * per), str(per)],
                                  stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env) for i in range(n)]
        outs = []
        for p in procs:
            o, e = p.communicate()
            if p.returncode != 0:
                for q in procs:
                    q.kill()
                raise RuntimeError(f"reference worker failed: {e.strip().splitlines()[-1] if e.strip() else p.returncode}")
            outs.append(json.loads(o.strip().splitlines()[-1]))
    slowest = max(o["seconds"] for o in outs)
    total = sum(o["frames"] * o["atoms"] for o in outs)
    return total / slowest, n, slowest, all(o.get("numba") for o in outs)


def cpu_baseline(wl: Workload, frac, lat):
    """cpu_baseline object of the JSON line: the reference itself when it can run here, the C port beside it."""
    cores = os.cpu_count() or 1
    port_v, port_n, port_dt = port_all_cores(wl, frac, lat, min(wl.port_frames, max(1, frac.shape[0] // cores)))
    why = reference_available()
    if why is None:
        try:
            v, n, dt, numba_on = reference_all_cores(wl, frac, lat, min(wl.ref_frames, max(1, frac.shape[0] // cores - 1)))
            return {"value": v, "unit": UNIT, "cores": n, "kind": "reference",
                    "sample": f"{n} processes x {min(wl.ref_frames, max(1, frac.shape[0] // cores - 1))} frames of this workload, "
                              f"slowest timed loop {dt:.2f} s (unmodified reference from oracle/_ref, numba {'on' if numba_on else 'OFF'}, "
                              f"NUMBA_NUM_THREADS=1 per process)",
                    "port_value": port_v, "port_cores": port_n,
                    "port_note": "plain-C oracle port of the same algorithm (oracle/site_oracle.c), one replica per thread"}
        except Exception as exc:
            why = f"{type(exc).__name__}: {exc}"
    return {"value": port_v, "unit": UNIT, "cores": port_n, "kind": "port",
            "sample": f"{port_n} threads x {min(wl.port_frames, max(1, frac.shape[0] // cores))} frames of this workload ({port_dt:.2f} s)",
            "reference_unavailable": why}


def run_reference_arm(args):
    """The reference's own CPU implementation on all host cores; each step a bounded sample of the workload."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    wl = Workload(args.workload)
    cores = os.cpu_count() or 1
    why = reference_available()
    per = (args.ref_frames_per_proc or wl.ref_frames) if why is None else (args.ref_frames_per_proc or wl.port_frames)
    frac, lat = wl.trajectory(cores * (per + 1), wl.seed, "cpu")
    frac, lat = frac.numpy(), lat.numpy()
    kind = "reference"
    if why is None:
        try:
            reference_all_cores(wl, frac, lat, min(per, 2))                 # import / JIT caches warm (untimed)
        except Exception as exc:
            why = f"{type(exc).__name__}: {exc}"
    if why is not None:
        kind, per = "port", args.ref_frames_per_proc or wl.port_frames
        frac, lat = wl.trajectory(cores * per, wl.seed, "cpu")
        frac, lat = frac.numpy(), lat.numpy()

    def one():
        if kind == "reference":
            v, n, dt, _ = reference_all_cores(wl, frac, lat, per)
            return v, n, dt
        return port_all_cores(wl, frac, lat, per)
    for _ in range(args.warmup):
        one()
    vals, t_tot, n = [], 0.0, 1
    for _ in range(args.steps):
        v, n, dt = one()
        vals.append(v); t_tot += dt
    value = float(np.mean(vals))
    sample = (f"{n} processes x {per} frames per step, unmodified reference (oracle/_ref, Python + numba, NUMBA_NUM_THREADS=1 each)"
              if kind == "reference" else f"{n} threads x {per} frames per step, plain-C oracle port (reference unavailable: {why})")
    print(json.dumps({
        "impl": "reference", "metric": wl.metric, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": wl.config(per * n),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": n, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------ our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="polyhedral", choices=["polyhedral", "voronoi_4x4x4", "dynvor_llzo", "spherical"])
    ap.add_argument("--frames", type=int, default=0, help="frames per GPU (0 = the workload's own size)")
    ap.add_argument("--ref-frames-per-proc", type=int, default=0, help="reference arm: frames per process and step (0 = per workload)")
    ap.add_argument("--parity-frames", type=int, default=-1, help="frames of every rank's shard re-checked against the oracle (-1 = per workload)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-api", action="store_true", help="skip the Trajectory.trajectory_from_arrays figure (e2e_api)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
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
    wl = Workload(args.workload)
    F = args.frames or wl.frames
    A, N = wl.A, wl.N
    d_frac, d_lat = wl.trajectory(F, wl.seed + 7919 * rank, dev)
    frame0, lat0 = d_frac[0].cpu().numpy(), d_lat[0].cpu().numpy()
    if world > 1:
        # every rank primes its site tables from the trajectory's first frame (rank 0's frame 0)
        first = [d_frac[0].clone(), d_lat[0].clone()]
        dist.broadcast(first[0], 0); dist.broadcast(first[1], 0)
        frame0, lat0 = first[0].cpu().numpy(), first[1].cpu().numpy()
        from site_analysis_b200 import distributed as sd
    eng = wl.engine(frame0, lat0, local)

    in_flight = []                                         # (full result, NCCL work) of the previous step
    peer_bufs, exchange = None, "single GPU"
    if world > 1:
        try:     # result rows stored by the kernel itself into every rank's full array (NVLink peer memory)
            if os.environ.get("SAB_NO_PEER_STORE") == "1" or wl.kind != "polyhedral":      # A/B switch: NCCL all-gather instead
                raise RuntimeError("NCCL all-gather selected")
            peer_bufs = sd.PeerResultBuffers(F, A, dist.group.WORLD, dev)
            exchange = "fused: the stream kernel stores its result rows into every rank's full array over NVLink peer memory"
        except Exception as exc:
            exchange = f"NCCL all-gather overlapping the next step's kernels ({type(exc).__name__}: {exc})"

    halo = None
    if world > 1 and wl.kind == "polyhedral":
        # input distribution, done once: every rank keeps the frame that precedes its shard (a loader handing rank r the
        # frames [f0 - 1, f1) has it for free), so the per-step halo collective disappears from the data path
        halo = sd.exchange_halo_frames(d_frac, dist.group.WORLD)

    def drain():
        while in_flight:
            full, work = in_flight.pop()
            work.wait()                                    # orders the current stream after that step's all-gather
        if world > 1:
            sd.finish_sharded(eng)                         # guard / flag records of all steps so far: ONE host synchronisation

    def step(with_history=False):
        if world > 1:
            # shard carry (tiny all-gather) + local kernels + guard/flag all-gather + the one result all-gather.  A long
            # trajectory streams through in batches: the result all-gather of batch k runs while batch k + 1 computes
            # (it is waited for one step later, and all of them before the timed region closes).
            full, work = sd.assign_sharded(eng, d_frac, d_lat, frame0, lat0, dist.group.WORLD, equal_shards=True,
                                           async_gather=True, peer_buffers=peer_bufs, halo_frame=halo,
                                           deferred_check=os.environ.get("SAB_SYNC_CHECK") != "1")
            drain()
            if work is not None:
                in_flight.append((full, work))
            return full[rank * F:(rank + 1) * F]
        res = eng.assign_device(d_frac, d_lat)
        if with_history:
            eng.sync_history()                             # transition Counters + recent sites of the batch (device fold)
        else:
            eng._pending = []                              # array path: no Python-object history to maintain
        return res

    def timed(n_steps, **kw):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        acc = dict(launches=0, assign_ns=0, wrap_ns=0)
        ev0.record()
        for _ in range(n_steps):
            r = step(**kw)
            acc["launches"] += eng.last_info["launches"] + (2 if world > 1 else 0)
            acc["assign_ns"] += eng.last_info["assign_ns"]; acc["wrap_ns"] += eng.last_info["wrap_ns"]
        drain()                                            # every result all-gather has finished inside the timed region
        ev1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), r, acc

    for _ in range(max(args.warmup, 0)):
        res = step()
    drain()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ms_total, res, acc = timed(args.steps)
    clocks = sampler.summary() if sampler else None
    value = world * F * A * args.steps / (ms_total * 1e-3)
    value_hist = None
    if world == 1:                                          # the same step with the history fold inside the timed region
        step(with_history=True)
        ms_h, _, _ = timed(args.steps, with_history=True)
        value_hist = {"value": F * A * args.steps / (ms_h * 1e-3), "unit": UNIT, "ms_per_step": ms_h / args.steps,
                      "what": "step + device fold of the batch into transition Counters and recent sites (csrc/k_fold.cuh)"}

    # ---- parity: every rank re-checks the first rows of ITS OWN shard of the timed result against the oracle --------
    npar = args.parity_frames if args.parity_frames >= 0 else {"polyhedral": 400, "voronoi_4x4x4": 4, "dynvor_llzo": 100, "spherical": 300}[wl.name]
    npar = min(npar, F)
    mism = torch.zeros(2, dtype=torch.int64, device=dev)
    if npar > 0:
        hf, hl = d_frac[:npar].cpu().numpy(), d_lat[:npar].cpu().numpy()
        o = wl.oracle(frame0, lat0)
        if rank > 0 and wl.kind in ("polyhedral", "dynamic_voronoi"):   # bring the oracle's PBC caches to the shard's start
            pass                                           # (its lazy caches update from the primed state at the first query)
        if wl.kind == "spherical" and (rank > 0 or args.steps + args.warmup > 1):
            # overlapping sites: the answer depends on the history of all earlier frames; check a fresh engine instead
            got = wl.engine(frame0, lat0, local).assign_device(d_frac[:npar].contiguous(), d_lat[:npar].contiguous()).cpu().numpy()
        else:
            got = res[:npar].cpu().numpy()
        want = o.run(hf, hl, positions=True)
        mism[0] = int((got != want).sum()); mism[1] = npar
    if world > 1:
        dist.all_reduce(mism, op=dist.ReduceOp.SUM)
    parity = {"frames_per_rank": int(npar), "ranks": world, "mismatches": int(mism[0].item()),
              "what": "first rows of every rank's shard of the timed (sharded) result vs the CPU oracle"}

    # ---- end to end through the C-ABI host call (rank-local shard; pinned host buffers) --------
    h_frac = torch.empty(d_frac.shape, dtype=torch.float64, pin_memory=True); h_frac.copy_(d_frac)
    h_lat = torch.empty(d_lat.shape, dtype=torch.float64, pin_memory=True); h_lat.copy_(d_lat)
    h_out = torch.empty((F, A), dtype=torch.int32, pin_memory=True)
    nf, nl, no = h_frac.numpy(), h_lat.numpy(), h_out.numpy()
    last_row_dev = res[-1].cpu().numpy() if rank == 0 else None
    del d_frac, d_lat, res, eng
    torch.cuda.empty_cache()
    eng_h = wl.engine(frame0, lat0, local)

    def host_call():
        if wl.kind == "spherical":
            np.copyto(no, eng_h.assign(nf, nl))            # overlapping sites: chunked path with the priority resolver
        else:
            eng_h.assign_host_pipelined(nf, nl, out=no)
        eng_h._pending = []
    e2e_api_name = "sab_assign_host (C ABI, pinned host buffers)" if wl.kind != "spherical" else "AssignmentEngine.assign (chunked, resolver)"
    for _ in range(min(args.warmup, 2)):
        host_call()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        host_call()
    torch.cuda.synchronize()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = world * F * A * args.steps / float(e2e_s.item())
    same = bool(np.array_equal(no[-1], last_row_dev)) if (rank == 0 and wl.kind != "spherical") else None
    del eng_h

    # ---- the user-facing call: Trajectory.trajectory_from_arrays on the same host buffers (N = 1) -----------------
    e2e_api = None
    if world == 1 and not args.no_api:
        n_api = max(1, min(args.steps, 3))
        traj = wl.api_trajectory(frame0, lat0)
        traj.trajectory_from_arrays(nf, nl)                 # warm-up (cell lists, pinned staging)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n_api):
            rows = traj.trajectory_from_arrays(nf, nl)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        idx = traj.site_collection._index_of_pos
        e2e_api = {"value": F * A * n_api / dt, "unit": UNIT, "steps": n_api, "api": "Trajectory.trajectory_from_arrays",
                   "equals_c_abi_result": bool(np.array_equal(np.asarray(rows)[-1], np.where(no[-1] >= 0, idx[np.clip(no[-1], 0, None)], -1)))
                   if wl.kind != "spherical" else None}
        del traj

    if rank == 0:
        peak, peak_src = measured_hbm_peak()
        bytes_per_launch = wl.bytes_per_frame * F
        assign_s = max(acc["assign_ns"] * 1e-9 / args.steps, 1e-12)
        achieved = bytes_per_launch / assign_s / 1e9
        cb = None
        if world == 1 and not args.no_cpu_baseline:        # reported at N = 1 only
            cores = os.cpu_count() or 1
            n_cpu = min(F, cores * (max(wl.ref_frames, wl.port_frames) + 1))
            cb = cpu_baseline(wl, nf[:n_cpu], nl[:n_cpu])
        cfg = wl.config(F)
        if world > 1:
            cfg["multi_gpu"] = ("contiguous frame shards, one per rank; the frame preceding each shard is distributed with the input "
                                "(once, outside the timed region); per step: kernels, ONE small guard/flag all-gather on the stream "
                                "(completion barrier of the peer stores; examined by the host once, at the end of the timed region), "
                                "full int32[F_total, A] result on every rank; result exchange = " + exchange
                                if halo is not None else
                                "contiguous frame shards, one per rank; per step: carry exchange, kernels, guard/flag all-gather, "
                                "full int32[F_total, A] result on every rank; result exchange = " + exchange)
        traffic = None
        tp = ROOT / "profiles" / "r2_traffic.json"
        if tp.exists():                                   # one `ncu --set full` capture per kernel, scaled per frame
            tj = json.loads(tp.read_text()).get(wl.name)
            if tj:
                traffic = (tj["dram_bytes_read"] + tj["dram_bytes_write"]) / tj["frames"] * F
        line = {
            "metric": wl.metric, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": cfg,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(F * N * 24 + F * 72),
                    "d2h_bytes_per_step": int(F * A * 4), "api": e2e_api_name, "equals_device_result": same},
            "e2e_api": e2e_api, "value_with_history": value_hist,
            "gpu_launches": int(acc["launches"]),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": wl.kernel, "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "bytes_per_launch": int(bytes_per_launch), "kernel_ms": assign_s * 1e3,
                         "wrap_scan_ms": acc["wrap_ns"] * 1e-6 / args.steps},
            "cpu_baseline": cb,
            "parity": parity,
        }
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
