#!/usr/bin/env python
"""bench.py — throughput of the voxelization hot path on B200 (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg4|cfg2|cfg3|cfg5] [--molecules M] [--chunk B]
                    [--augment] [--gather] [--out-dtype float32|bfloat16|float16]
    python bench.py --impl reference ...        # the reference's CPU path on the host cores

Default workload = BASELINE.json's metric configuration (cfg4): the virtual-screening SWEEP of 1,000,000 DISTINCT
synthetic ligands (40-60 atoms, 1.5 A random walk; generated on the device by a counter-based generator keyed by the
global molecule index, so any sharding sees the same molecules), forward_types, 9 channels, 64^3, Gaussian.  The sweep
is sharded over the ranks by molecule index (strong scaling), every rank walks its shard in chunks of 1,024 molecules
through a ring of two output buffers (9.7 GB per chunk >> 126 MB L2), and the K timed "steps" are K equal parts of the
shard: the timed region is the whole sweep whatever --steps is.  The other workloads (cfg2, cfg3, cfg5) cycle through a
pool of distinct synthetic batches; their steps are sized so that the timed region lasts >= 1 s.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: mode, C, dim, res, density, radii_type, atoms/molecule, molecules per call (chunk) per GPU
    "cfg4": dict(mode="types", C=9, dim=64, res=0.5, density="gaussian", radii_type="scalar", atoms=(40, 60), batch=1024,
                 molecules=1_000_000,
                 desc="virtual-screening sweep: synthetic ligands (40-60 atoms, 1.5 A random walk), forward_types 9 channels, 64^3, res 0.5, gaussian sigma 0.5, r 1.0"),
    "cfg3": dict(mode="types", C=4, dim=64, res=0.5, density="binary", radii_type="scalar", atoms=(40, 60), batch=1024,
                 molecules=262_144,
                 desc="binary-density forward_types 4 channels, 64^3, batches of 1,024 ligands (~50 atoms)"),
    "cfg2": dict(mode="features", C=16, dim=48, res=0.5, density="gaussian", radii_type="scalar", atoms=(2000, 2000), batch=256,
                 pool=8,
                 desc="synthetic protein pocket 2,000 atoms, forward_features C=16 (8 one-hot + 8 Bernoulli(0.25)), 48^3, gaussian, batch 256"),
    "cfg2b": dict(mode="features", C=16, dim=48, res=0.5, density="gaussian", radii_type="scalar", atoms=(2000, 2000), batch=256,
                  pool=8,
                  desc="cfg2 with the atoms in a Gaussian blob (sigma 4 A) instead of uniform: heterogeneous tiles (not a BASELINE config)"),
    "cfg5": dict(mode="features", C=32, dim=96, res=0.375, density="gaussian", radii_type="atom-wise", atoms=(10000, 10000), batch=32,
                 pool=4,
                 desc="large complex 10,000 atoms, forward_features C=32 dense, 96^3, res 0.375, atom-wise radii U[1,2]"),
}
SWEEP_SEED = 4   # cfg4 uses seed 4

ATOMS_OVERRIDE = 0


def make_batch(name: str, B: int, seed: int):
    """Seeded synthetic inputs of SURVEY.md §8d on the HOST (coordinates rounded to fp32-representable values): the
    pool batches of the dense workloads and the CPU legs' samples."""
    w = WORKLOADS[name]
    rng = np.random.default_rng(seed)
    lo, hi = w["atoms"]
    uniform_cube = lo >= 1000   # pocket / complex workloads; ligands are random walks
    if ATOMS_OVERRIDE:          # density sweeps (--atoms): same distribution, another atom count
        lo = hi = ATOMS_OVERRIDE
    counts = rng.integers(lo, hi + 1, size=B)
    offs = np.zeros(B + 1, dtype=np.int32)
    offs[1:] = np.cumsum(counts)
    N = int(offs[-1])
    mol_of = np.repeat(np.arange(B), counts)
    if not uniform_cube:   # ligand: 3-D random walk with 1.5 A steps, recentred
        d = rng.normal(size=(N, 3))
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        d *= 1.5
        cs = np.cumsum(d, axis=0)
        start = cs[offs[:-1]] - d[offs[:-1]]
        xyz = cs - start[mol_of]
        mean = np.add.reduceat(xyz, offs[:-1], axis=0) / counts[:, None]
        xyz -= mean[mol_of]
    else:           # pocket / complex: uniform in the grid cube
        half = w["res"] * (w["dim"] - 1) / 2.0
        xyz = rng.uniform(-half, half, size=(N, 3))
        if name == "cfg2b":
            xyz = np.clip(rng.normal(scale=4.0, size=(N, 3)), -half, half)
    coords = xyz.astype(np.float32).astype(np.float64)
    out = dict(offs=offs, coords=coords, centers=np.zeros((B, 3)), types=None, feats=None, radii=1.0)
    if w["mode"] == "types":
        out["types"] = rng.integers(0, w["C"], size=N).astype(np.int32)
    elif name in ("cfg2", "cfg2b"):
        f = np.zeros((N, 16), dtype=np.float32)
        f[np.arange(N), rng.integers(0, 8, size=N)] = 1.0
        f[:, 8:] = (rng.uniform(size=(N, 8)) < 0.25).astype(np.float32)
        out["feats"] = f
    else:
        out["feats"] = rng.uniform(0, 1, size=(N, w["C"])).astype(np.float32)
    if w["radii_type"] == "atom-wise":
        out["radii"] = rng.uniform(1.0, 2.0, size=N).astype(np.float32)
    return out


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def is_sweep(name: str) -> bool:
    return "molecules" in WORKLOADS[name]


def workload_config(args, world: int) -> dict:
    """The `config` object of the JSON line: identical for both arms (the reference arm times a bounded sample of it)."""
    w = WORKLOADS[args.workload]
    cfg = {"workload": f"{args.workload}: {w['desc']}", "channels": w["C"], "grid": f"{w['dim']}^3", "resolution": w["res"],
           "density": w["density"], "radii_type": w["radii_type"], "chunk": args.batch or w["batch"],
           "augment": bool(args.augment), "compat_blockdim": 8, "out_dtype": args.out_dtype}
    if getattr(args, "channels_last", False):
        cfg["out_layout"] = "channels-last (B, D, H, W, C) — not the headline (reference) layout"
    if is_sweep(args.workload):
        cfg["molecules"] = args.molecules or w["molecules"]
        cfg["inputs"] = "distinct molecules from a counter-based generator keyed by the global molecule index (mvx_synth_ligands)"
    else:
        cfg["pool_batches"] = w["pool"]
        cfg["inputs"] = "pool of distinct seeded batches, cycled"
    return cfg


# --------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation on the host cores (never touches CUDA)
# --------------------------------------------------------------------------------------------
_REF_STATE = {}


def _ref_init(name, library):
    sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
    import molvoxel
    w = WORKLOADS[name]
    _REF_STATE["vox"] = molvoxel.create_voxelizer(w["res"], w["dim"], w["radii_type"], w["density"], library=library)
    _REF_STATE["w"] = w


def _ref_run(job):
    coords, channels, radii, rt, rr = job
    vox, w = _REF_STATE["vox"], _REF_STATE["w"]
    center = np.zeros(3)
    if w["mode"] == "types":
        g = vox.forward_types(coords, center, channels.astype(np.int16), radii, rt, rr)
    else:
        g = vox.forward_features(coords, center, channels, radii, rt, rr)
    return float(g[0, 0, 0, 0])


def _jobs(batch, w, augment=False):
    offs = batch["offs"]
    ch = batch["types"] if w["mode"] == "types" else batch["feats"]
    jobs = []
    for m in range(len(offs) - 1):
        a, b = offs[m], offs[m + 1]
        r = batch["radii"] if np.isscalar(batch["radii"]) else batch["radii"][a:b]
        jobs.append((batch["coords"][a:b], ch[a:b], r, 0.5 if augment else 0.0, bool(augment)))
    return jobs


def reference_available() -> bool:
    return os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "molvoxel"))


def time_reference(name, steps, warmup, library="numpy", procs=None, per_core=None, augment=False):
    """Molecules/s of the unmodified reference (baseline/_ref) over a Pool of independent instances."""
    import multiprocessing as mp
    w = WORKLOADS[name]
    cores = procs or host_cores()
    if per_core is None:
        per_core = {"cfg4": 64, "cfg3": 64, "cfg2": 8, "cfg2b": 8, "cfg5": 1}[name]
    sample = cores * per_core
    jobs = _jobs(make_batch(name, sample, seed=1234), w, augment)
    ctx = mp.get_context("fork")
    with ctx.Pool(cores, initializer=_ref_init, initargs=(name, library)) as pool:
        chunk = max(1, per_core // 4)
        for _ in range(warmup):
            pool.map(_ref_run, jobs, chunksize=chunk)
        t0 = time.perf_counter()
        for _ in range(steps):
            pool.map(_ref_run, jobs, chunksize=chunk)
        dt = time.perf_counter() - t0
    return dict(value=sample * steps / dt, ms_per_step=dt / steps * 1e3, cores=cores, kind="reference",
                sample=f"{sample} molecules/step ({per_core} per core) of the {name} workload, library={library!r}, "
                       f"{cores} forked reference instances")


def time_reference_single_core(name, library, n):
    _ref_init(name, library)
    jobs = _jobs(make_batch(name, n, seed=1234), WORKLOADS[name])
    _ref_run(jobs[0])   # numba: JIT warm-up excluded
    t0 = time.perf_counter()
    for j in jobs:
        _ref_run(j)
    return n / (time.perf_counter() - t0)


def time_oracle_port(name, steps, warmup, per_core=None):
    """Molecules/s of the oracle (C restatement) with all host threads — used when baseline/_ref is absent."""
    from oracle import oracle_forward_batch
    w = WORKLOADS[name]
    cores = host_cores()
    if per_core is None:
        per_core = {"cfg4": 16, "cfg3": 16, "cfg2": 4, "cfg2b": 4, "cfg5": 1}[name]
    sample = min(cores * per_core, max(cores, int(6e9 // (4 * w["C"] * w["dim"] ** 3))))
    b = make_batch(name, sample, seed=1234)
    run = lambda: oracle_forward_batch(w["res"], w["dim"], w["radii_type"], w["density"], 0.5, 8, w["mode"], b["offs"],  # noqa: E731
                                       b["coords"], b["centers"], b["types"], b["feats"], w["C"], b["radii"],
                                       num_threads=cores)
    for _ in range(warmup):
        run()
    t0 = time.perf_counter()
    for _ in range(steps):
        run()
    dt = time.perf_counter() - t0
    return dict(value=sample * steps / dt, ms_per_step=dt / steps * 1e3, cores=cores, kind="port",
                sample=f"{sample} molecules/step of the {name} workload, oracle/mvx_oracle.c on {cores} pthreads")


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if reference_available():
        r = time_reference(args.workload, args.steps, args.warmup, "numpy", augment=args.augment)
    else:
        r = time_oracle_port(args.workload, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": "molecules_per_sec", "value": r["value"], "unit": "molecules/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "strong" if is_sweep(args.workload) else "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(args, int(os.environ.get("WORLD_SIZE", "1"))),
        "note": "reference CPU path (numpy backend, stock forward_*) on the host cores; a step is a bounded sample of the workload",
        "cpu_baseline": {"value": r["value"], "unit": "molecules/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": "molecules/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "25", "-i", str(index)], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in ln.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1] or [r for _, r in self.rows[-3:]]
        sm, mx, reasons, power = [], None, set(), []
        for r in rows:
            try:
                sm.append(float(r[0])); mx = float(r[1]); power.append(float(r[2]))
            except Exception:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power) if power else None}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy bandwidth)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(name, batch):
    """DRAM bytes of one voxelize launch from the committed ncu capture; the capture's batch is recorded beside it and
    the figure is only reported for that batch."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return d.get(name) if int(d.get("batch", {}).get(name, -1)) == int(batch) else None
        except Exception:
            return None
    return None


class Chunks:
    """One rank's share of the workload as a list of CSR chunks, device-resident (for `value`) and in pinned host
    memory (for `e2e`).  chunk(i) -> dict of per-call arguments; first_mol(i) = global index of its first molecule."""

    def __init__(self, name, args, rank, world, dev):
        import torch
        import molvoxel_b200 as mv
        from molvoxel_b200 import _lib
        self.torch = torch
        self.w = w = WORKLOADS[name]
        self.B = B = args.batch or w["batch"]
        self.dev = dev
        self.sweep = is_sweep(name) and not ATOMS_OVERRIDE
        self.keep = []
        if self.sweep:
            M = args.molecules or w["molecules"]
            self.lo, self.hi = mv.shard_bounds(M, rank, world)
            n_mol = self.hi - self.lo
            L = _lib.lib()
            st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            counts = torch.empty(max(n_mol, 1), dtype=torch.int32, device=dev)
            vmin, vmax = w["atoms"]
            _lib.raise_for_status(L.mvx_synth_ligands(SWEEP_SEED, self.lo, n_mol, vmin, vmax, w["C"], 1.5, None,
                                                      ctypes.c_void_p(counts.data_ptr()), None, _lib.MVX_F64, None, st))
            offs = torch.zeros(n_mol + 1, dtype=torch.int32, device=dev)
            offs[1:] = torch.cumsum(counts[:n_mol], 0)
            N = int(offs[-1])
            self.coords = torch.empty((N, 3), dtype=torch.float64, device=dev)
            self.types = torch.empty(N, dtype=torch.int32, device=dev)
            _lib.raise_for_status(L.mvx_synth_ligands(SWEEP_SEED, self.lo, n_mol, vmin, vmax, w["C"], 1.5,
                                                      ctypes.c_void_p(offs.data_ptr()), None, ctypes.c_void_p(self.coords.data_ptr()),
                                                      _lib.MVX_F64, ctypes.c_void_p(self.types.data_ptr()), st))
            torch.cuda.synchronize()
            self.offs_h = offs.cpu().numpy()
            self.n = (n_mol + B - 1) // B
            # per-chunk CSR offsets (rebased to 0), resident on the device and in pinned host memory
            self.bounds = [(c * B, min(n_mol, (c + 1) * B)) for c in range(self.n)]
            self.offs_chunks_h = [np.ascontiguousarray(self.offs_h[m0:m1 + 1] - self.offs_h[m0]).astype(np.int32) for m0, m1 in self.bounds]
            self.offs_chunks_d = [torch.from_numpy(o).to(dev) for o in self.offs_chunks_h]
            self.centers_d = torch.zeros((B, 3), dtype=torch.float64, device=dev)
            self.n_mols = n_mol
            self.atoms = N
            self.radii, self.max_r = 1.0, None
            self.host = None
        else:
            P = w.get("pool", 8)
            self.n = P
            self.batches = [make_batch(name, B, seed=1000 + 97 * rank + p) for p in range(P)]
            self.dev_batches = []
            for b in self.batches:
                ch = b["types"] if w["mode"] == "types" else b["feats"]
                self.dev_batches.append(dict(
                    offs=torch.from_numpy(b["offs"]).to(dev), coords=torch.from_numpy(b["coords"]).to(dev),
                    centers=torch.from_numpy(b["centers"]).to(dev), chan=torch.from_numpy(ch).to(dev),
                    radii=b["radii"] if np.isscalar(b["radii"]) else torch.from_numpy(b["radii"]).to(dev)))
            self.max_r = None if np.isscalar(self.batches[0]["radii"]) else float(max(b["radii"].max() for b in self.batches))
            self.n_mols = B * P
            self.atoms = int(sum(int(b["offs"][-1]) for b in self.batches))
            self.lo = rank * (1 << 32)
            self.host = None

    def mols(self, i):
        if self.sweep:
            m0, m1 = self.bounds[i]
            return m1 - m0
        return self.B

    def first_mol(self, i):
        return self.lo + (self.bounds[i][0] if self.sweep else i * self.B)

    def device_args(self, i):
        if self.sweep:
            m0, m1 = self.bounds[i]
            a0, a1 = int(self.offs_h[m0]), int(self.offs_h[m1])
            return dict(coords=self.coords[a0:a1], offs=self.offs_chunks_d[i], centers=self.centers_d[:m1 - m0],
                        chan=self.types[a0:a1], radii=1.0)
        return self.dev_batches[i]

    def pin_host(self):
        """Pinned host copies of the inputs (made once, before the e2e timed region)."""
        torch = self.torch
        if self.host is not None:
            return

        def pin(a):
            t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
            self.keep.append(t)
            return t.numpy()
        self.host = []
        if self.sweep:
            coords_h = pin(self.coords.cpu().numpy())
            types_h = pin(self.types.cpu().numpy())
            centers_h = pin(np.zeros((self.B, 3)))
            for i, (m0, m1) in enumerate(self.bounds):
                a0, a1 = int(self.offs_h[m0]), int(self.offs_h[m1])
                self.host.append(dict(coords=coords_h[a0:a1], offs=pin(self.offs_chunks_h[i]), centers=centers_h[:m1 - m0],
                                      chan=types_h[a0:a1], radii=1.0))
        else:
            # the loader's job: per-molecule point clouds -> pinned CSR batch.  compact=True picks the
            # lossless small encodings by itself (uint8 rows for 0/1 features, float32 for fp32-representable coordinates
            # next to float64 centres); dense float features / fp64-only coordinates stay as they are.
            from molvoxel_b200.pointcloud import Collator, PointCloud
            self.host_plain = []
            for b in self.batches:
                chn = b["types"] if self.w["mode"] == "types" else b["feats"]
                offs = b["offs"]
                clouds = [PointCloud(b["coords"][offs[m]:offs[m + 1]], chn[offs[m]:offs[m + 1]], self.w["C"]) for m in range(self.B)]
                rad = None if np.isscalar(b["radii"]) else [b["radii"][offs[m]:offs[m + 1]] for m in range(self.B)]
                for compact, dst in ((True, self.host), (False, self.host_plain)):
                    col = Collator(pinned=True, compact=compact)
                    self.keep.append(col)
                    c = col(clouds, centers=b["centers"], radii=rad)
                    dst.append(dict(coords=c["coords"], offs=c["mol_offsets"], centers=c["centers"], chan=c["channels"],
                                    radii=b["radii"] if np.isscalar(b["radii"]) else c["radii"]))

    def host_args(self, i):
        return self.host[i]

    def host_bytes(self, i):
        return sum(int(v.nbytes) for v in self.host[i].values() if isinstance(v, np.ndarray))


def run_b200_arm(args):
    import torch
    import torch.distributed as dist

    import molvoxel_b200 as mv
    from molvoxel_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    name = args.workload
    w = WORKLOADS[name]
    K, W = args.steps, max(3, args.warmup)
    D, C = w["dim"], w["C"]
    out_dt = getattr(torch, args.out_dtype)
    esize = 4 if args.out_dtype == "float32" else 2
    vox = mv.create_voxelizer(w["res"], D, w["radii_type"], w["density"], library="b200", device=dev, out_dtype=out_dt,
                              seed=SWEEP_SEED, overlap=not args.no_overlap, channels_last=args.channels_last)
    ch = Chunks(name, args, rank, world, dev)
    B = ch.B
    ring = [vox.get_empty_grid(C, B) for _ in range(2)]   # (B, C, D, D, D); physically (B, D, D, D, C) with --channels-last
    rt, rr = (0.5, True) if args.augment else (0.0, False)
    fwd = {"types": lambda a, **kw: vox.forward_types_batch(a["coords"], a["offs"], a["centers"], a["chan"], a["radii"], C, rt, rr, **kw),
           "features": lambda a, **kw: vox.forward_features_batch(a["coords"], a["offs"], a["centers"], a["chan"], a["radii"], rt, rr, **kw)}[w["mode"]]

    # device-resident sweep inputs are complete before the timed region: the library may run a call's prep / binning on its
    # second stream next to the previous call's voxelize kernel (inputs_ready; ligand kernels only)
    ready = {"inputs_ready": True} if w["mode"] == "types" and not args.no_overlap else {}

    def call_device(i, k):
        fwd(ch.device_args(i), out=ring[k & 1][:ch.mols(i)], max_radius=ch.max_r, rng_offset=ch.first_mol(i), **ready)

    def call_host(i, k):   # public API, host inputs, pipelined: H2D of call k+1 overlaps the kernels of call k
        fwd(ch.host_args(i), out=ring[k & 1][:ch.mols(i)], max_radius=ch.max_r, rng_offset=ch.first_mol(i), non_blocking=True)

    def call_host_blocking(i, k):   # mvx_voxelize_host: one synchronisation per call
        fwd(ch.host_args(i), out=ring[k & 1][:ch.mols(i)], max_radius=ch.max_r, rng_offset=ch.first_mol(i))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- step plan: which chunk-calls make up each of the K timed steps --------------------------------
    for k in range(3):   # library warm-up + a first time estimate
        call_device(k % ch.n, k)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(4):
        call_device(k % ch.n, k)
    torch.cuda.synchronize()
    est_call_s = (time.perf_counter() - t0) / 4
    if ch.sweep:   # K equal parts of the shard: the timed region is the whole sweep
        edges = np.linspace(0, ch.n, K + 1).round().astype(int)
        steps = [list(range(edges[s], edges[s + 1])) for s in range(K)]
        warm_steps = [list(range(min(ch.n, max(1, ch.n // K))))] * W
    else:          # pool: cycle, with enough calls per step for a timed region of >= args.min_seconds
        per_step = max(1, int(np.ceil(args.min_seconds / max(est_call_s, 1e-6) / K)))
        steps = [[(s * per_step + j) % ch.n for j in range(per_step)] for s in range(K)]
        warm_steps = steps[:1] * W
    n_calls = sum(len(s) for s in steps)
    n_mols_timed = sum(ch.mols(i) for s in steps for i in s)

    def timed(call, step_list, warm, profile=False):
        k = 0
        for s in warm:
            for i in s:
                call(i, k); k += 1
        barrier()
        if profile:
            _lib.raise_for_status(_lib.lib().mvx_profile_begin(sum(len(s) for s in step_list)))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for s in step_list:
            for i in s:
                call(i, k); k += 1
        e1.record()
        barrier()
        t1 = time.perf_counter()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        prof = None
        if profile:
            a, b_, c = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
            n = ctypes.c_int()
            _lib.raise_for_status(_lib.lib().mvx_profile_end(ctypes.byref(a), ctypes.byref(b_), ctypes.byref(c), ctypes.byref(n)))
            prof = dict(prep=a.value / max(1, n.value), bin=b_.value / max(1, n.value), vox=c.value / max(1, n.value), calls=n.value)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms), prof, (t0, t1)

    # context for the roofline: a pure write stream (torch fill kernel) over the same ring buffers
    ms_fill, _, _ = timed(lambda i, k: ring[k & 1].zero_(), [[0] * 50], [[0] * 3])
    fill_gbs = B * float(esize) * C * D ** 3 * 50 / (ms_fill * 1e-3) / 1e9

    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms_total, prof, (tw0, tw1) = timed(call_device, steps, warm_steps, profile=True)
    clocks = sampler.stop(tw0, tw1) if sampler else None
    vox.check_status()

    ch.pin_host()
    ms_e2e, _, _ = timed(call_host, steps, warm_steps)
    vox.check_status()
    nb = min(n_calls, 64)
    flat = [i for s in steps for i in s][:nb]
    ms_e2e_blocking, _, _ = timed(call_host_blocking, [flat], [flat[:3]])
    ms_e2e_blocking *= n_calls / nb
    h2d = float(np.mean([ch.host_bytes(i) for s in steps for i in s])) * (n_calls / K)

    tot = torch.tensor([float(n_mols_timed)], device=dev)
    if world > 1:
        dist.all_reduce(tot)
    mols_all = float(tot)

    # ---- extras (reported, never fail the bench) --------------------------------------------------------
    extras = {}
    if not ch.sweep:
        extras["plain_inputs"] = plain_inputs_leg(ch, vox, fwd, ring, timed, steps, warm_steps, world, K, n_calls)
    if not args.channels_last:
        extras["with_grid_d2h"] = grid_d2h_leg(ch, vox, fwd, ring, esize, C, D, out_dt, world, dev)
    if esize == 4 and not args.channels_last:
        extras["with_grid_d2h_sparse"] = grid_d2h_sparse_leg(ch, vox, fwd, ring, C, D, world, dev)
    if args.gather and world > 1:
        extras["gather"] = gather_leg(ring[0], world, dev)

    spec = vox._spec()
    a0 = ch.device_args(0)
    bb = _lib.Batch()
    bb.mode, bb.num_mols, bb.total_atoms = _lib.MODE[w["mode"]], ch.mols(0), int(a0["coords"].shape[0])
    bb.num_channels = bb.out_channels = C
    dummy = ctypes.c_void_p(256)
    bb.mol_offsets = bb.coords = bb.types = bb.features = bb.radii = dummy
    bb.radius, bb.max_radius = 1.0, 2.0
    per_call = _lib.lib().mvx_launches_per_call(ctypes.byref(spec), ctypes.byref(bb))
    kernel_name = _lib.FORM_KERNEL.get(_lib.lib().mvx_voxelize_form(ctypes.byref(spec), ctypes.byref(bb)), "mvx_voxelize_kernel")

    parity = parity_sample(ch, vox, w, C, D, rt, rr) if rank == 0 and not args.no_parity else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = mols_all / (ms_total * 1e-3)
    e2e_value = mols_all / (ms_e2e * 1e-3)
    out_bytes = float(esize) * C * D ** 3
    atoms_per_mol = ch.atoms / max(1, ch.n_mols)
    in_bytes = atoms_per_mol * (3 * 8 + (4 if w["mode"] == "types" else 4 * C) + (4 if w["radii_type"] == "atom-wise" else 0))
    mols_per_call = n_mols_timed / n_calls
    alg_bytes = mols_per_call * (out_bytes + in_bytes)
    peak, peak_src = measured_peak()
    achieved = alg_bytes / (prof["vox"] * 1e-3) / 1e9
    cfg = workload_config(args, world)
    cfg.update({"parallelism": f"dp{world} (molecules sharded by index, no data-path collective)",
                "l2_policy": "every call writes >= 1.8 GB of grids into a ring of 2 buffers, far beyond the 126 MB L2; inputs are distinct per call; no flush needed",
                "overlap_binning": bool(ready) and vox._overlap is not None,
                "calls_per_step_per_gpu": n_calls / K, "molecules_per_call": mols_per_call,
                "atoms_per_molecule": atoms_per_mol, "out_bytes_per_call_per_gpu": int(mols_per_call * out_bytes)})
    line = {
        "metric": "molecules_per_sec", "value": value, "unit": "molecules/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "strong" if ch.sweep else "weak", "vs_baseline": None,
        "dtype": "f32" if esize == 4 else f"f32 compute, {args.out_dtype} output (not the headline metric)", "data": "synthetic",
        "config": cfg,
        "timed_region_s": ms_total * 1e-3, "molecules_timed": int(mols_all),
        "e2e": {"value": e2e_value, "unit": "molecules/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(4 * n_calls / K),
                "ms_per_step": ms_e2e / K, "blocking_value": mols_all / (ms_e2e_blocking * 1e-3), **extras,
                "input_encoding": {k: str(v.dtype) for k, v in ch.host_args(0).items() if isinstance(v, np.ndarray)},
                "note": "public API with pinned HOST inputs (pool workloads: collated by molvoxel_b200.Collator(pinned=True, compact=True), which narrows losslessly where it can), every call: Voxelizer.forward_*_batch(non_blocking=True) = async H2D (copy stream, 2-deep staging ring) -> prep/bin/voxelize -> D2H of the status word; one sync at the end of the timed region. blocking_value = mvx_voxelize_host (one sync per call). Grids stay in HBM (reference torch-backend convention); with_grid_d2h* = grids delivered to host memory"},
        "gpu_launches": per_call * n_calls,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": ncu_traffic(name, B), "kernel": kernel_name, "kernel_ms": prof["vox"],
                     "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                     "write_only_reference_gbs": fill_gbs,
                     "note": "peak is the measured COPY bandwidth (read+write); this kernel only writes, so frac can exceed 1.0 — write_only_reference_gbs is torch zero_() (device memset) on the same buffers",
                     "step_share": {"prep_ms": prof["prep"], "bin_ms": prof["bin"], "voxelize_ms": prof["vox"], "calls": prof["calls"],
                                    "note": "per call; bin_ms = column/layer binning + entry build kernels between prep and voxelize"}},
        "clocks": clocks, "parity": parity,
    }
    if world == 1:   # the reference's own calling pattern: one molecule per call, host arrays in (cfg 1 shape)
        line["single_call"] = single_call_leg(mv, w, D, C, dev, out_dt, name)
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_subprocess(name)
        try:
            line["cpu_baseline"].setdefault("detail", {})["reference_torch_cuda"] = reference_torch_cuda_leg(name, dev)
        except Exception as e:
            line["cpu_baseline"].setdefault("detail", {})["reference_torch_cuda"] = {"error": repr(e)[:300]}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def plain_inputs_leg(ch, vox, fwd, ring, timed, steps, warm_steps, world, K, n_calls):
    """The same e2e path with the host batches collated WITHOUT the lossless compaction (float64 coordinates, float32
    feature rows): what `e2e` was before the loader learned to pick the small encodings.  Identical grids."""
    try:
        same = all(h["chan"].dtype == p["chan"].dtype and h["coords"].dtype == p["coords"].dtype for h, p in zip(ch.host, ch.host_plain))
        if same:
            return {"note": "the collator found nothing to narrow for this workload; e2e already uses these inputs"}

        def call(i, k):
            fwd(ch.host_plain[i], out=ring[k & 1][:ch.mols(i)], max_radius=ch.max_r, rng_offset=ch.first_mol(i), non_blocking=True)
        ms, _, _ = timed(call, steps, warm_steps)
        vox.check_status()
        nbytes = float(np.mean([sum(int(v.nbytes) for v in c.values() if isinstance(v, np.ndarray)) for c in ch.host_plain]))
        return {"value": world * sum(ch.mols(i) for s in steps for i in s) / (ms * 1e-3), "h2d_bytes_per_step": int(nbytes * n_calls / K),
                "note": "float64 coordinates + float32 feature rows over PCIe (Collator(compact=False))"}
    except Exception as e:
        return {"error": repr(e)[:200]}


def grid_d2h_leg(ch, vox, fwd, ring, esize, C, D, out_dt, world, dev):
    """The pipelined host path with the finished grids also copied back to pinned HOST memory every call (what a
    host-side consumer would see).  Bounded: a slice of the chunk, a few calls — it measures PCIe."""
    import torch
    try:
        ch.pin_host()
        B = ch.mols(0)
        Bd = max(1, min(B, int(2e9 // (esize * C * D ** 3))))
        Kd = 8
        h0 = ch.host_args(0)
        offs = np.ascontiguousarray(h0["offs"][:Bd + 1])
        nd = int(offs[-1])
        small = dict(coords=h0["coords"][:nd], offs=offs, centers=h0["centers"][:Bd], chan=h0["chan"][:nd],
                     radii=h0["radii"] if np.isscalar(h0["radii"]) else h0["radii"][:nd])
        host_grid = torch.empty((2, Bd, C, D, D, D), dtype=out_dt).pin_memory()
        d2h_stream = torch.cuda.Stream(dev)
        done = [torch.cuda.Event(), torch.cuda.Event()]

        def step(k):
            done[k & 1].synchronize()   # the host slot of call k-2 has been read back
            fwd(small, out=ring[k & 1][:Bd], max_radius=ch.max_r, rng_offset=ch.first_mol(0), non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            with torch.cuda.stream(d2h_stream):
                d2h_stream.wait_event(ev)
                host_grid[k & 1].copy_(ring[k & 1][:Bd], non_blocking=True)
                done[k & 1].record(d2h_stream)
        for k in range(2):
            step(k)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k in range(Kd):
            step(k)
        torch.cuda.synchronize()
        sec = time.perf_counter() - t0
        vox.check_status()
        return {"value": world * Bd * Kd / sec, "batch": Bd, "calls": Kd, "d2h_bytes_per_call": int(Bd * esize * C * D ** 3),
                "note": "dense grids copied to pinned host memory every call (separate D2H stream, 2 host slots); bounded slice, wall clock — PCIe-bound"}
    except Exception as e:
        return {"error": repr(e)[:200]}


def grid_d2h_sparse_leg(ch, vox, fwd, ring, C, D, world, dev):
    """Grids delivered to HOST memory in brick-sparse form (Voxelizer.compact_into: non-empty 8x8x8 bricks + ids): the
    pipelined host path, then the compaction kernel on the same stream, the brick count read back, and the payload of
    exactly that many bricks copied to pinned host memory on a D2H stream, two calls in flight.  Bounded: sub-chunks of
    <= 256 molecules, 24 calls; one sub-chunk is rebuilt on the host (SparseGrids.to_dense) and compared bit for bit."""
    import torch
    from molvoxel_b200.sparse import SparseGrids
    try:
        ch.pin_host()
        Bs = max(1, min(ch.mols(0), 256, int(2e9 // (4 * C * D ** 3))))
        nb = -(-D // 8)
        cap = max(4096, Bs * C * nb ** 3 // (4 if ch.sweep else 1))
        picks = list(range(min(ch.n, 24)))
        subs = []
        for i in picks:
            h = ch.host_args(i)
            offs = np.ascontiguousarray(h["offs"][:Bs + 1])
            na = int(offs[-1])
            subs.append(dict(coords=h["coords"][:na], offs=offs, centers=h["centers"][:Bs], chan=h["chan"][:na],
                             radii=h["radii"] if np.isscalar(h["radii"]) else h["radii"][:na]))
        slots = [dict(ids=torch.empty(cap, dtype=torch.int32, device=dev), vals=torch.empty((cap, 512), dtype=torch.float32, device=dev),
                      count=torch.zeros(1, dtype=torch.int32, device=dev), h_count=torch.zeros(1, dtype=torch.int32).pin_memory(),
                      h_ids=torch.empty(cap, dtype=torch.int32).pin_memory(), h_vals=torch.empty((cap, 512), dtype=torch.float32).pin_memory(),
                      counted=torch.cuda.Event(), copied=torch.cuda.Event(), n=0) for _ in range(2)]
        d2h = torch.cuda.Stream(dev)
        cur = torch.cuda.current_stream(dev)

        def issue(k):
            s = slots[k & 1]
            s["copied"].synchronize()          # the host slot of call k-2 has been filled (and may be consumed)
            cur.wait_event(s["copied"])        # ... and its device bricks are no longer being read
            out = ring[k & 1][:Bs]
            fwd(subs[k % len(subs)], out=out, max_radius=ch.max_r, rng_offset=ch.first_mol(picks[k % len(subs)]), non_blocking=True)
            vox.compact_into(out, s["ids"], s["vals"], s["count"])
            s["h_count"].copy_(s["count"], non_blocking=True)
            s["counted"].record(cur)

        def drain(k):
            s = slots[k & 1]
            s["counted"].synchronize()
            n = int(s["h_count"][0])
            if n > cap:
                raise RuntimeError(f"brick capacity {cap} too small for {n}")
            with torch.cuda.stream(d2h):
                d2h.wait_event(s["counted"])
                s["h_ids"][:n].copy_(s["ids"][:n], non_blocking=True)
                s["h_vals"][:n].copy_(s["vals"][:n], non_blocking=True)
                s["copied"].record(d2h)
            s["n"] = n
            return n

        def run(K):
            total = 0
            for k in range(K):
                issue(k)
                if k >= 1:
                    total += drain(k - 1)
            total += drain(K - 1)
            torch.cuda.synchronize()
            return total
        run(4)
        K = 24
        t0 = time.perf_counter()
        bricks = run(K)
        sec = time.perf_counter() - t0
        vox.check_status()
        # exactness of what landed on the host: the last call, rebuilt densely
        s = slots[(K - 1) & 1]
        sp = SparseGrids(s["h_ids"][:s["n"]].numpy(), s["h_vals"][:s["n"]].numpy(), (Bs, C, D))
        exact = bool(np.array_equal(sp.to_dense(), ring[(K - 1) & 1][:Bs].cpu().numpy()))
        dense_bytes = 4.0 * C * D ** 3
        return {"value": world * Bs * K / sec, "batch": Bs, "calls": K, "bricks_per_molecule": bricks / (Bs * K),
                "d2h_bytes_per_molecule": bricks / (Bs * K) * 2052.0, "dense_bytes_per_molecule": dense_bytes,
                "host_rebuild_bit_exact": exact,
                "note": "non-empty 8x8x8 bricks + ids copied to pinned host memory (mvx_compact_bricks after every call, 2 calls in flight); bounded, wall clock"}
    except Exception as e:
        return {"error": repr(e)[:300]}


def gather_leg(grids, world, dev):
    """Optional NCCL all-gather of finished grids over NVLink (off the measured path): every rank contributes a
    slice of one output buffer; GB/s received per GPU."""
    import torch
    import torch.distributed as dist
    try:
        per = max(1, min(grids.shape[0], int(1.0e9 // (grids[0].numel() * grids.element_size()))))   # ~1 GB per rank
        src = grids[:per].contiguous()
        out = torch.empty((world * per,) + tuple(src.shape[1:]), dtype=src.dtype, device=dev)
        for _ in range(3):
            dist.all_gather_into_tensor(out, src)
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 10
        e0.record()
        for _ in range(n):
            dist.all_gather_into_tensor(out, src)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        nbytes = src.numel() * src.element_size()
        return {"bytes_per_rank": int(nbytes), "ms": float(ms), "recv_gbs_per_gpu": (world - 1) * nbytes / (float(ms) * 1e-3) / 1e9,
                "molecules_per_s_gathered": world * per / (float(ms) * 1e-3),
                "note": "torch.distributed.all_gather_into_tensor (NCCL) of finished grids, max over ranks; NVLink-5 nominal 900 GB/s per direction"}
    except Exception as e:
        return {"error": repr(e)[:200]}


def parity_sample(ch, vox, w, C, D, rt, rr):
    """Parity inside the run: >= 1,024 molecules of this rank's share (8 chunks spread over the shard, 128 molecules
    each; all of a dense pool batch) re-voxelized and compared with the oracle on the SAME inputs (device-generated
    inputs are copied back; device-drawn transforms are fetched and applied on the host with the reference's arithmetic)."""
    import torch
    from oracle import oracle_forward_batch
    from molvoxel_b200.transform import do_transform
    try:
        per = 128 if ch.sweep else max(1, min(ch.B, int(1.5e9 // (4 * C * D ** 3))))
        picks = sorted(set(np.linspace(0, ch.n - 1, 8 if ch.sweep else min(ch.n, 4)).round().astype(int).tolist()))
        worst, nmol, support_ok, exact = 0.0, 0, True, True
        for i in picks:
            a = ch.device_args(i)
            m = min(per, ch.mols(i))
            offs = a["offs"][:m + 1].cpu().numpy()
            na = int(offs[-1])
            coords = a["coords"][:na].cpu().numpy()
            chan = a["chan"][:na].cpu().numpy()
            radii = a["radii"] if np.isscalar(a["radii"]) else a["radii"][:na].cpu().numpy()
            sub = dict(coords=a["coords"][:na], offs=a["offs"][:m + 1], centers=a["centers"][:m], chan=a["chan"][:na],
                       radii=a["radii"] if np.isscalar(a["radii"]) else a["radii"][:na])
            kw = dict(max_radius=ch.max_r, rng_offset=ch.first_mol(i))
            if w["mode"] == "types":
                got = vox.forward_types_batch(sub["coords"], sub["offs"], sub["centers"], sub["chan"], sub["radii"], C, rt, rr, **kw)
            else:
                got = vox.forward_features_batch(sub["coords"], sub["offs"], sub["centers"], sub["chan"], sub["radii"], rt, rr, **kw)
            got = got.cpu().numpy()
            if rt or rr:
                rows = vox.random_transforms(m, rt, rr, rng_offset=ch.first_mol(i)).cpu().numpy()
                coords = np.concatenate([do_transform(coords[offs[j]:offs[j + 1]], None, rows[j, 4:].astype(np.float32).reshape(1, 3),
                                                      tuple(rows[j, :4])) for j in range(m)])
            ref = oracle_forward_batch(w["res"], D, w["radii_type"], w["density"], 0.5, 8, w["mode"], offs, coords, np.zeros((m, 3)),
                                       chan if w["mode"] == "types" else None, chan if w["mode"] == "features" else None, C, radii,
                                       num_threads=host_cores())
            peak = max(1.0, float(np.abs(ref).max()))
            worst = max(worst, float(np.abs(got - ref).max()) / peak)
            support_ok = support_ok and bool(np.array_equal(got != 0, ref != 0))
            exact = exact and bool(np.array_equal(got, ref))
            nmol += m
        return {"molecules": nmol, "max_abs_err_over_peak": worst, "support_identical": support_ok, "bit_exact": exact,
                "tolerance": 0.0 if w["density"] == "binary" else 1e-5,
                "ok": bool(exact if w["density"] == "binary" and w["mode"] == "types" else (support_ok and worst <= 1e-5)),
                "note": "CUDA path vs oracle/mvx_oracle.c on the same inputs, sampled across this rank's share after the timed region"}
    except Exception as e:
        return {"error": repr(e)[:300]}


def single_call_leg(mv, w, D, C, dev, out_dt, name):
    import torch
    b = make_batch(name, 1, seed=5)
    one = mv.create_voxelizer(w["res"], D, w["radii_type"], w["density"], library="b200", device=dev, out_dtype=out_dt)
    grid1 = one.get_empty_grid(C)
    ch1 = b["types"] if w["mode"] == "types" else b["feats"]
    lat = []
    for _ in range(60):
        t0 = time.perf_counter()
        one.forward(b["coords"], b["centers"][0], ch1, b["radii"], out_grid=grid1)
        torch.cuda.synchronize()
        lat.append(time.perf_counter() - t0)
    return {"median_us": float(np.median(lat[10:]) * 1e6), "atoms": int(b["offs"][-1]),
            "note": "Voxelizer.forward on one molecule with numpy inputs, synchronised (reference calling pattern)"}


def reference_torch_cuda_leg(name, dev):
    """The reference's only GPU path: its torch backend with device='cuda' (torch/voxelizer.py:255-330, :539-567), stock
    code from baseline/_ref, on the BINARY twin of the workload (its Gaussian cutoff is broken upstream)."""
    import torch
    if not reference_available():
        return {"unavailable": "baseline/_ref missing"}
    sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
    import molvoxel
    w = WORKLOADS[name]
    n = {"cfg4": 48, "cfg3": 48, "cfg2": 6, "cfg2b": 6, "cfg5": 2}[name]
    b = make_batch(name, n, seed=1234)
    vox = molvoxel.create_voxelizer(w["res"], w["dim"], w["radii_type"], "binary", library="torch", device=str(dev))
    jobs = []
    for m in range(n):
        a0, a1 = int(b["offs"][m]), int(b["offs"][m + 1])
        coords = torch.from_numpy(b["coords"][a0:a1]).to(dev, torch.float32)
        chn = torch.from_numpy((b["types"] if w["mode"] == "types" else b["feats"])[a0:a1]).to(dev)
        r = b["radii"] if np.isscalar(b["radii"]) else torch.from_numpy(b["radii"][a0:a1]).to(dev)
        jobs.append((coords, chn.long() if w["mode"] == "types" else chn, r))
    center = torch.zeros(3, device=dev)
    grid = vox.get_empty_grid(w["C"])

    def run(j):
        coords, chn, r = j
        if w["mode"] == "types":
            vox.forward_types(coords, center, chn, r, out_grid=grid)
        else:
            vox.forward_features(coords, center, chn, r, out_grid=grid)
    for j in jobs[:3]:
        run(j)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for j in jobs:
        run(j)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "molecules/s", "sample": f"{n} molecules, one stock forward_* call each, inputs resident on the GPU, binary density",
            "note": "unmodified reference torch backend on this B200 (stock ATen ops, 8 blocks per grid)"}


def cpu_baseline_subprocess(name):
    """Run the CPU legs in a fresh process (no CUDA context to fork)."""
    try:
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--cpu-baseline-worker", "--workload", name],
                             capture_output=True, text=True, timeout=900)
        for ln in reversed(out.stdout.strip().splitlines()):
            if ln.startswith("{"):
                return json.loads(ln)
        return {"error": (out.stderr or out.stdout)[-400:]}
    except Exception as e:   # the baseline is a reported number; never fail the bench because of it
        return {"error": repr(e)}


def cpu_baseline_worker(name):
    res = {}
    detail = {}
    n1 = {"cfg4": 64, "cfg3": 64, "cfg2": 6, "cfg2b": 6, "cfg5": 2}[name]
    if reference_available():
        res = time_reference(name, steps=4, warmup=1, library="numpy")
        detail["numpy_1core_mol_per_s"] = time_reference_single_core(name, "numpy", n1)
        try:
            detail["numba_1core_mol_per_s"] = time_reference_single_core(name, "numba", n1)
            detail["numba_allcores_mol_per_s"] = time_reference(name, steps=2, warmup=1, library="numba")["value"]
        except Exception as e:
            detail["numba_error"] = repr(e)[:200]
        import numpy, scipy
        detail["versions"] = {"numpy": numpy.__version__, "scipy": scipy.__version__}
    port = time_oracle_port(name, steps=2, warmup=1)
    detail["oracle_port_allcores_mol_per_s"] = port["value"]
    if not res:
        res = port
    out = {"value": res["value"], "unit": "molecules/s", "cores": res["cores"], "kind": res["kind"], "sample": res["sample"],
           "detail": detail}
    try:
        out["cpu_model"] = [ln.split(":", 1)[1].strip() for ln in open("/proc/cpuinfo") if ln.startswith("model name")][0]
    except Exception:
        pass
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=0, help="timed steps (default 20); the timed region is the whole sweep / >= --min-seconds whatever K is")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=sorted(WORKLOADS))
    ap.add_argument("--molecules", type=int, default=0, help="sweep workloads: total distinct molecules over all GPUs (default: the workload's, 1,000,000 for cfg4)")
    ap.add_argument("--batch", "--chunk", type=int, default=0, dest="batch", help="molecules per call per GPU")
    ap.add_argument("--min-seconds", type=float, default=1.0, help="pool workloads: lower bound of the timed region")
    ap.add_argument("--augment", action="store_true", help="random_translation=0.5, random_rotation=True (the reference's batched use-case), drawn on the device")
    ap.add_argument("--gather", action="store_true", help="N > 1: also time the optional NCCL all-gather of finished grids")
    ap.add_argument("--no-overlap", action="store_true", help="keep every call's prep / binning on the caller's stream (A/B of mvx_voxelize_split)")
    ap.add_argument("--channels-last", action="store_true", help="grids in the channels-last layout (B, D, H, W, C) (SURVEY row f3; not the headline layout)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--atoms", type=int, default=0, help="atoms per molecule (density sweeps; default: the workload's own)")
    ap.add_argument("--out-dtype", default="float32", choices=["float32", "bfloat16", "float16"],
                    help="reduced-precision output grids (not the headline metric, which is fp32)")
    ap.add_argument("--cpu-baseline-worker", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    global ATOMS_OVERRIDE
    ATOMS_OVERRIDE = max(0, args.atoms)
    if args.steps <= 0:
        args.steps = 20 if args.impl == "b200" else 5
    if args.cpu_baseline_worker:
        return cpu_baseline_worker(args.workload)
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_b200_arm(args)


if __name__ == "__main__":
    main()
