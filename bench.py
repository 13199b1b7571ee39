#!/usr/bin/env python
"""bench.py -- grid-point Laplacian steps/s of the B200 filter path vs the HBM roofline.

    python bench.py --gpus N --steps K --warmup W [--workload cfg3] [--impl reference]

One "step" = one pass of the hot path (a whole filter call: prepare, n_steps Chebyshev steps,
finalize) over one synthetic batch.  Default workload (BASELINE.json north_star headline, SURVEY
8(d) cfg3): Gaussian IRREGULAR_WITH_LAND filter, filter_scale 36 / dx_min 0.9 -> n_steps 44, on a
62 x 2400 x 3600 fp64 POP 0.1-degree field (NaN on land).  Units are grid-point Laplacian steps
(points x n_steps).

  value     device-resident throughput: inputs already in HBM, CUDA events around K filter calls
  e2e       the same through the public API Filter.apply() with pinned HOST buffers (H2D + D2H inside)
  roofline  dominant kernel (the mid-recurrence step kernel): algorithmic bytes (B_alg = 5w*ncomp + C/nb
            per pt-step) / CUDA-event duration per launch, vs MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the numpy oracle port (oracle/np_oracle.py) on the host cores, bounded sample

Multi-GPU (torchrun, one rank per GPU): the batch dimension shards with no data-path collective;
every rank filters its own 62-level field (weak scaling); NCCL is used for the barrier and the
max-over-ranks of the timings only.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "grid_point_laplacian_steps_per_sec"
UNIT = "pt-steps/s"


# ------------------------------------------------------------------------------------------ workloads
def build_workload(name, nb, rank=0):
    import bench_inputs as fixtures  # deterministic synthetic inputs (no oracle, no product code)

    if name == "cfg3":
        cfg = fixtures.cfg3(nb=nb or 62)
    elif name == "cfg3f32":
        cfg = fixtures.cfg3(nb=nb or 62, dtype=np.float32)
    elif name == "cfg3taper":
        cfg = fixtures.cfg3(nb=nb or 62, gaussian=False)
    elif name == "cfg2":
        cfg = fixtures.cfg2(nb=nb or 365)
    elif name == "cfg4":
        cfg = fixtures.cfg4(nb=nb or None)
    elif name == "cfg5":
        cfg = fixtures.cfg5()
    elif name == "cfg1":
        cfg = fixtures.cfg1()
    else:
        raise SystemExit(f"unknown workload {name}")
    if rank:  # weak scaling: every rank filters a different field of the same shape
        cfg["fields"] = tuple(f * f.dtype.type(1.0 + 0.01 * rank) for f in cfg["fields"])
    return cfg


def c_bytes_per_point(grid_type, w):
    """Coefficient-plane bytes per grid point (C in B_alg = 5 w ncomp + C/nb)."""
    return {"REGULAR": 0, "REGULAR_AREA_WEIGHTED": 0, "REGULAR_WITH_LAND": 1, "REGULAR_WITH_LAND_AREA_WEIGHTED": 1,
            "TRIPOLAR_REGULAR_WITH_LAND_AREA_WEIGHTED": 1, "IRREGULAR_WITH_LAND": 3 * w, "MOM5U": 3 * w,
            "MOM5T": 3 * w, "TRIPOLAR_POP_WITH_LAND": 3 * w, "VECTOR_B_GRID": 8 * w, "VECTOR_C_GRID": 14 * w}[grid_type]


def workload_descr(cfg, n_steps):
    f0 = cfg["fields"][0]
    return {"workload": f"{cfg['name']}: {cfg['grid_type']} {cfg['filter_args']['filter_shape']} "
                        f"filter_scale={cfg['filter_args']['filter_scale']:g} dx_min={cfg['filter_args']['dx_min']:g} "
                        f"n_steps={n_steps} field={'x'.join(str(s) for s in f0.shape)} {f0.dtype}",
            "grid_type": cfg["grid_type"], "n_steps": int(n_steps), "shape": list(f0.shape),
            "l2_policy": "inputs larger than L2 (no flush needed)" if f0.nbytes > 2 * 126e6 else
                         "L2 flushed between timed iterations"}


# ------------------------------------------------------------------------------------------ CPU arm
_CPU = {}


def _cpu_worker(args):
    """One worker = one 2-D slice through the oracle's recurrence (numpy, single-threaded ufuncs)."""
    from oracle import np_oracle

    idx, n_steps = args
    cfg = _CPU["cfg"]
    fa = dict(cfg["filter_args"], n_steps=n_steps)
    fields = tuple(f[idx] if f.ndim == 3 else f for f in cfg["fields"])
    t0 = time.perf_counter()
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        np_oracle.apply_filter(cfg["grid_type"], cfg["grid_vars"], fields, **fa)
    return time.perf_counter() - t0


def cpu_baseline(cfg, n_steps_sample=8, max_workers=32, repeats=1):
    """Oracle port on the host cores: `workers` processes x 1 slice each x n_steps_sample steps."""
    import multiprocessing as mp

    cores = len(os.sched_getaffinity(0))
    f0 = cfg["fields"][0]
    nslices = f0.shape[0] if f0.ndim == 3 else 1
    workers = max(1, min(cores, max_workers, nslices))
    _CPU["cfg"] = cfg
    pts = f0.shape[-1] * f0.shape[-2]
    best = None
    ctx = mp.get_context("fork")
    with ctx.Pool(workers) as pool:
        for _ in range(repeats):
            t0 = time.perf_counter()
            pool.map(_cpu_worker, [(i, n_steps_sample) for i in range(workers)])
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    value = workers * pts * n_steps_sample / best
    return {"value": value, "unit": UNIT, "cores": workers, "kind": "port",
            "sample": f"{workers} slice(s) of {f0.shape[-2]}x{f0.shape[-1]} ({cfg['grid_type']}, {f0.dtype}) x "
                      f"n_steps={n_steps_sample} (per-step cost is constant), one process per slice, "
                      f"numpy oracle port; host has {cores} usable cores", "seconds": best}


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            pass

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        out, _ = self.proc.communicate(timeout=10)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------ GPU arm
def measured_hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def gpu_arm(args):
    import torch
    import torch.distributed as dist

    from gcm_filters_b200 import Filter, FilterShape, GridType, _cabi, engine
    from gcm_filters_b200.filter import _shift_scale

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun for --gpus > 1")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    engine.set_steps_per_block(args.steps_per_block)
    cfg = build_workload(args.workload, args.nb, rank)
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        cpu = cpu_baseline(cfg)  # before any CUDA work in this process' children (fork)

    fa = dict(cfg["filter_args"])
    fa["filter_shape"] = FilterShape[fa["filter_shape"]]
    flt = Filter(grid_type=GridType[cfg["grid_type"]], grid_vars=cfg["grid_vars"], **fa)
    n_steps = int(flt.n_steps)
    lap = flt.laplacian
    lib = _cabi.get_library()
    spec = flt.filter_spec
    c = _shift_scale(spec, lap)

    fields = cfg["fields"]
    shape = fields[0].shape
    ny, nx = shape[-2:]
    nb = int(np.prod(shape[:-2])) if len(shape) > 2 else 1
    w = fields[0].dtype.itemsize
    ncomp = len(fields)
    units_per_step = nb * ny * nx * n_steps  # grid-point Laplacian steps in one filter call

    # host buffers (pinned) for the end-to-end leg, device-resident copies for the kernel leg
    host_in = [torch.from_numpy(np.ascontiguousarray(f)).pin_memory() for f in fields]
    host_out = [torch.empty_like(h).pin_memory() for h in host_in]
    dev_in = [h.to(dev).reshape(nb, ny, nx) for h in host_in]
    dev_out = [torch.empty_like(d) for d in dev_in]
    small = dev_in[0].numel() * w <= 2 * 126e6
    flush = torch.empty(int(300e6), dtype=torch.uint8, device=dev) if small else None

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    stream = torch.cuda.current_stream(dev)

    # ---- leg 1: device-resident filter calls -------------------------------------------------
    for _ in range(args.warmup):
        engine.filter_device(lap, spec.p, c, dev_in, dev_out)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = lib.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for k in range(args.steps):
        if flush is not None:
            flush.fill_(k & 0xFF)
        ev[k][0].record(stream)
        engine.filter_device(lap, spec.p, c, dev_in, dev_out)
        ev[k][1].record(stream)
    barrier()
    launches = lib.launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    dev_ms = max_over_ranks(dev_ms)
    value = world * units_per_step * args.steps / (dev_ms * 1e-3)

    # ---- leg 2: end to end through Filter.apply with pinned host buffers ------------------------
    def e2e_once():
        if ncomp == 1:
            flt.apply(host_in[0], dims=["y", "x"], out=host_out[0])
        else:
            flt.apply_to_vector(host_in[0], host_in[1], dims=["y", "x"], out=host_out)

    for _ in range(max(1, min(args.warmup, 2))):
        e2e_once()
    barrier()
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_once()  # returns after the D2H copy of the result has completed
    torch.cuda.synchronize(dev)
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = world * units_per_step * e2e_steps / e2e_s
    h2d = sum(h.numel() * h.element_size() for h in host_in)
    d2h = sum(h.numel() * h.element_size() for h in host_out)

    # ---- leg 3: per-launch durations, same launch sequence as gcmf_filter, CUDA events per launch ----
    plan = engine.device_plan(lap, local, fields[0].dtype, ny, nx)
    plan.set_filter(spec.p, c)
    wsb = lib.workspace_bytes(plan.handle, nb)
    ws = engine.workspace(dev, wsb)
    kfuse = lib.fused_max_steps(plan.handle) if engine.STEPS_PER_BLOCK != 1 else 0
    nbuf = 4 if kfuse else 2
    bufbytes = wsb // (nbuf * ncomp)
    spec_of = lambda ts: [(t.data_ptr(), nx, ny * nx) for t in ts]
    bufs = [[(ws.data_ptr() + (j * ncomp + k) * bufbytes, nx, ny * nx) for k in range(ncomp)] for j in range(nbuf)]
    A, B = bufs[0], bufs[1]
    sptr = stream.cuda_stream
    records = []
    area = cfg["grid_type"] in ("REGULAR_AREA_WEIGHTED", "REGULAR_WITH_LAND_AREA_WEIGHTED",
                                "TRIPOLAR_REGULAR_WITH_LAND_AREA_WEIGHTED")

    def timed(kind, k, fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        records.append((kind, k, e0, e1))

    outs = spec_of(dev_out)
    for rep in range(max(1, min(args.steps, 3))):
        X = spec_of(dev_in)
        if area:
            lib.prepare(plan.handle, nb, X, B, sptr)
            X = B
        if kfuse:  # same blocks as gcmf_filter: ceil(n/kfuse) fused launches, first/last steps included
            cur = 1 if area else 0
            T1, T2 = X, X
            i = 1
            while i <= n_steps:
                kk = min(n_steps - i + 1, kfuse)
                O1, O2 = bufs[2 * cur], bufs[2 * cur + 1]
                kind = "fused_first" if i == 1 else ("fused_last" if i + kk - 1 == n_steps else "fused")
                timed(kind, kk, lambda: lib.cheb_fused(plan.handle, nb, i, kk, T1, T2, O1, O2, outs, sptr))
                T1, T2 = O1, O2
                cur ^= 1
                i += kk
            continue
        timed("first", 1, lambda: lib.cheb_step(plan.handle, nb, 1, X, None, A, outs, sptr))
        T1, T2 = A, X
        for i in range(2, n_steps + 1):
            D = B if (i == 2 and not area) else T2
            timed("mid" if i < n_steps else "last", 1,
                  lambda: lib.cheb_step(plan.handle, nb, i, T1, T2, D, outs, sptr))
            T2, T1 = T1, D
    torch.cuda.synchronize(dev)
    per_kind = {}
    for kind, k, e0, e1 in records:
        d = per_kind.setdefault((kind, k), [])
        d.append(e0.elapsed_time(e1))
    dom = max(per_kind, key=lambda key: sum(per_kind[key]))
    dom_ms = float(np.mean(per_kind[dom]))
    b_alg = 5 * w * ncomp + c_bytes_per_point(cfg["grid_type"], w) / nb
    peak, peak_src = measured_hbm_peak()
    # algorithmic bytes of one launch = B_alg per grid-point step x points x steps the launch performs
    achieved = b_alg * nb * ny * nx * dom[1] / (dom_ms * 1e-3) / 1e9
    fk = f"fused_kernel ({dom[1]} Chebyshev steps per launch, TMA-staged tiles)"
    kname = {"fused": fk, "fused_first": fk + ", first block", "fused_last": fk + ", last block",
             "mid": "step_kernel<MODE_MID> (one Chebyshev step)", "first": "step_kernel<MODE_FIRST>",
             "last": "step_kernel<MODE_LAST>"}[dom[0]]
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "kernel": kname, "ms_per_launch": dom_ms, "steps_per_launch": dom[1],
                "algorithmic_bytes_per_pt_step": b_alg, "peak_source": peak_src,
                "launch_mix_ms": {f"{k[0]}x{k[1]}": [len(v), float(np.mean(v))] for k, v in per_kind.items()}}
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(prof):  # ncu-measured DRAM bytes per unit of the dominant kernel (committed capture)
        try:
            with open(prof) as fh:
                tr = json.load(fh)
            key = args.workload if dom[0].startswith("fused") else args.workload + "_onestep"
            if key in tr:
                roofline["traffic"] = tr[key]["bytes_per_pt_step"] * nb * ny * nx * dom[1]
                roofline["traffic_source"] = tr[key]["source"]
        except Exception:
            pass

    total_launches = int(sum_over_ranks(launches))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64" if w == 8 else "f32", "data": "synthetic (numpy PCG64, SURVEY 8(d))",
        "config": dict(workload_descr(cfg, n_steps), sharding="batch slabs, one full field per GPU, no collective"),
        "roofline": roofline, "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "api": "Filter.apply(pinned host tensor, out=pinned host tensor)"},
        "gpu_launches": total_launches, "clocks": clocks,
    }
    emit(line)


# ------------------------------------------------------------------------------------------ banded arm
def banded_arm(args):
    """One 2-D slice split into latitude bands, one band per GPU, ghost rows exchanged by NCCL point-to-point
    after every Chebyshev step (BASELINE config 5).  Strong scaling: the total work is fixed."""
    import torch
    import torch.distributed as dist

    from gcm_filters_b200 import Filter, FilterShape, GridType, _cabi
    from gcm_filters_b200.scheduler import BandedFilter, FusedBandedFilter, PeerBandedFilter

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = build_workload(args.workload, args.nb)
    fa = dict(cfg["filter_args"])
    fa["filter_shape"] = FilterShape[fa["filter_shape"]]
    flt = Filter(grid_type=GridType[cfg["grid_type"]], grid_vars=cfg["grid_vars"], **fa)
    n_steps = int(flt.n_steps)
    if args.fused:
        bf = FusedBandedFilter(flt, rank, world, exchange="peer" if args.peer else "nccl")
    else:
        bf = (PeerBandedFilter if args.peer else BandedFilter)(flt, rank, world)
    st = bf.stage(*cfg["fields"])
    f0 = cfg["fields"][0]
    ny, nx = f0.shape[-2:]
    nb = st["nb"]
    w = f0.dtype.itemsize
    lib = _cabi.get_library()

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        bf.run(st)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    n0 = lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stream = torch.cuda.current_stream(dev)
    e0.record(stream)
    for _ in range(args.steps):
        bf.run(st)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = lib.launch_count() - n0
    if world > 1:
        t = torch.tensor([ms, float(launches)], dtype=torch.float64, device=dev)
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        ms, launches = float(tmax[0].item()), int(t[1].item())
        dist.barrier()
        if hasattr(bf, "close"):
            st = None
            bf.close()
        dist.destroy_process_group()
    clocks = sampler.stop() if sampler else None
    if rank != 0:
        return
    units = nb * ny * nx * n_steps
    value = units * args.steps / (ms * 1e-3)
    ncomp = len(cfg["fields"])
    b_alg = 5 * w * ncomp + c_bytes_per_point(cfg["grid_type"], w) / nb
    peak, peak_src = measured_hbm_peak()
    achieved = b_alg * value / 1e9 / world  # per GPU
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64" if w == 8 else "f32", "data": "synthetic (numpy PCG64, SURVEY 8(d))",
        "config": dict(workload_descr(cfg, n_steps),
                       sharding=f"{world} latitude band(s), " +
                       ("4 ghost rows, fused 4-step blocks, ghost rows pulled from peer memory once per block" if args.fused and args.peer
                        else "4 ghost rows, fused 4-step blocks, one NCCL exchange per block" if args.fused else
                        "ghost rows stored by the step kernels into the neighbours' peer memory (NVLink), flag-synchronised"
                        if args.peer else "NCCL send/recv per Chebyshev step")),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "kernel": "step_kernel (one Chebyshev step per launch) + halo exchange, per GPU",
                     "algorithmic_bytes_per_pt_step": b_alg, "peak_source": peak_src},
        "cpu_baseline": None, "e2e": None, "gpu_launches": int(launches), "clocks": clocks,
    }
    emit(line)


# ------------------------------------------------------------------------------------------ reference arm
def reference_arm(args):
    """The reference is pure Python/numpy and cannot travel to the GPU box; its CPU implementation of
    the path is timed through the oracle port (bit-identical to it, tests/test_oracle.py) on all host
    cores, on a bounded sample of the same workload per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = build_workload(args.workload, args.nb)
    from oracle import np_oracle
    fa = cfg["filter_args"]
    n_steps = np_oracle.resolve_n_steps(fa["filter_scale"], fa["dx_min"], fa["filter_shape"])
    for _ in range(min(args.warmup, 1)):
        cpu_baseline(cfg, n_steps_sample=3)
    t_all, res = 0.0, None
    steps = max(1, min(args.steps, 3))
    for _ in range(steps):
        res = cpu_baseline(cfg)
        t_all += res["seconds"]
    f0 = cfg["fields"][0]
    units = res["cores"] * f0.shape[-1] * f0.shape[-2] * 8
    value = units * steps / t_all
    res["value"] = value
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * t_all / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64" if f0.dtype.itemsize == 8 else "f32",
        "data": "synthetic (numpy PCG64, SURVEY 8(d))", "config": workload_descr(cfg, n_steps),
        "cpu_baseline": res,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


class _CleanStdout:
    """Everything written to fd 1 while the benchmark runs (NCCL prints its version there, libraries may
    chatter) is diverted to stderr; the ONE JSON line goes to the real stdout."""

    def __enter__(self):
        sys.stdout.flush()
        self.real = os.dup(1)
        os.dup2(2, 1)
        return self

    def emit(self, text):
        os.write(self.real, (text + "\n").encode())

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.real, 1)
        os.close(self.real)
        return False


_OUT = None


def emit(line):
    text = json.dumps(line)
    if _OUT is not None:
        _OUT.emit(text)
    else:
        print(text)


def main():
    global _OUT
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg3")
    ap.add_argument("--nb", type=int, default=0, help="override the batch size (levels / time steps)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--steps-per-block", type=int, default=0, help="0 auto, 1 = one-step kernels only")
    ap.add_argument("--fused", action="store_true",
                    help="with --banded: temporally blocked kernel on the bands, one ghost exchange per 4-step block")
    ap.add_argument("--peer", action="store_true",
                    help="with --banded: ghost rows pushed by the step kernels through peer memory instead of NCCL")
    ap.add_argument("--banded", action="store_true",
                    help="latitude-band domain decomposition with NCCL halo exchange (strong scaling; cfg5)")
    args = ap.parse_args()
    with _CleanStdout() as out:
        _OUT = out
        try:
            if args.impl == "reference":
                reference_arm(args)
            elif args.banded:
                banded_arm(args)
            else:
                gpu_arm(args)
        finally:
            _OUT = None


if __name__ == "__main__":
    main()
