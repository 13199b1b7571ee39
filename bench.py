#!/usr/bin/env python
"""bench.py -- grid-point Laplacian steps/s of the B200 filter path vs the HBM roofline.

    python bench.py --gpus N --steps K --warmup W [--workload cfg3] [--impl reference]

One "step" = one pass of the hot path (a whole filter call: prepare, n_steps Chebyshev steps, finalize) over ONE
synthetic field.  Default workload (BASELINE.json north_star headline): Gaussian
IRREGULAR_WITH_LAND filter, filter_scale 36 / dx_min 0.9 -> n_steps 44, on a 62 x 2400 x 3600 fp64 POP 0.1-degree
field (NaN on land).  Units are grid-point Laplacian steps (points x n_steps).

  value        device-resident throughput: the field already in HBM, CUDA events around K filter calls
  e2e          the same through the public API Filter.apply() with pinned HOST buffers (H2D + D2H inside)
  e2e_numpy    the reference's own call signature: pageable numpy in -> numpy out
  roofline     dominant kernel (the mid-recurrence block): algorithmic bytes (B_alg = 5w*ncomp + C/nb per pt-step,
               DESIGN.md) / CUDA-event duration per launch, vs MEASURED_PEAKS.json hbm_gbs
  secondary    the other BASELINE configs under the same clock (cfg2, cfg4, cfg5, cfg1)
  cpu_baseline the reference's numpy path on the host cores, bounded sample (N = 1, rank 0)

Multi-GPU (torchrun, one rank per GPU): STRONG scaling of the one field -- its batch dimension (62 depth levels) is
cut into contiguous slabs by gcm_filters_b200.scheduler.batch_slabs (8,8,8,8,8,8,7,7 on 8 GPUs), every rank filters
its slab, no data-path collective; NCCL carries the barrier and the max-over-ranks of the timings only.  value = units
of the WHOLE field / max-over-ranks time.  The end-to-end leg scatters from / gathers into one host array (a shared
/dev/shm segment; every rank pins and first-touches its own slab).
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
DATA = "synthetic (numpy PCG64, SURVEY 8(d))"


# ------------------------------------------------------------------------------------------ workloads
def workload_nb(name, nb):
    """Batch size (levels / time steps) of the whole field of a workload; None = a single 2-D field."""
    default = {"cfg3": 62, "cfg3f32": 62, "cfg3taper": 62, "cfg2": 365, "cfg4": None, "cfg5": None, "cfg1": None,
               "cfgb": None}
    if name not in default:
        raise SystemExit(f"unknown workload {name}")
    if name in ("cfg5", "cfg1", "cfgb"):
        return None
    return nb or default[name]


def build_workload(name, nb=0, levels=None):
    """Deterministic synthetic inputs (no oracle, no product code).  ``levels=(a, b)``: only batch slices a..b-1 of
    the whole field -- the slab one rank of a batch-sharded run owns (same values as in the whole array)."""
    import bench_inputs as fixtures

    nbt = workload_nb(name, nb)
    if name == "cfg3":
        return fixtures.cfg3(nb=nbt, levels=levels)
    if name == "cfg3f32":
        return fixtures.cfg3(nb=nbt, dtype=np.float32, levels=levels)
    if name == "cfg3taper":
        return fixtures.cfg3(nb=nbt, gaussian=False, levels=levels)
    if name == "cfg2":
        return fixtures.cfg2(nb=nbt, levels=levels)
    if name == "cfg4":
        return fixtures.cfg4(nb=nbt, levels=levels if nbt else None)
    if name == "cfg5":
        return fixtures.cfg5()
    if name == "cfgb":
        return fixtures.cfgb()
    return fixtures.cfg1()


def c_bytes_per_point(grid_type, w):
    """Coefficient-plane bytes per grid point (C in B_alg = 5 w ncomp + C/nb)."""
    return {"REGULAR": 0, "REGULAR_AREA_WEIGHTED": 0, "REGULAR_WITH_LAND": 1, "REGULAR_WITH_LAND_AREA_WEIGHTED": 1,
            "TRIPOLAR_REGULAR_WITH_LAND_AREA_WEIGHTED": 1, "IRREGULAR_WITH_LAND": 3 * w, "MOM5U": 3 * w,
            "MOM5T": 3 * w, "TRIPOLAR_POP_WITH_LAND": 3 * w, "VECTOR_B_GRID": 8 * w, "VECTOR_C_GRID": 14 * w}[grid_type]


def config_of(cfg, n_steps, shape, dtype):
    """The workload description; identical for both arms (`--impl reference` times the same config)."""
    fa = cfg["filter_args"]
    nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
    return {"workload": f"{cfg['name']}: {cfg['grid_type']} {fa['filter_shape']} filter_scale={fa['filter_scale']:g} "
                        f"dx_min={fa['dx_min']:g} n_steps={n_steps} field={'x'.join(str(s) for s in shape)} "
                        f"{np.dtype(dtype).name}",
            "grid_type": cfg["grid_type"], "n_steps": int(n_steps), "shape": [int(s) for s in shape],
            "l2_policy": "inputs larger than L2 (no flush needed)" if nbytes > 2 * 126e6 else
                         "L2 flushed between timed iterations",
            "sharding": "ONE field; batch dimension cut into contiguous slabs, one per GPU (strong scaling), no "
                        "data-path collective"}


# ------------------------------------------------------------------------------------------ CPU arm
_CPU = {}


def _cpu_worker(args):
    """One worker = one 2-D slice through the reference's recurrence (numpy, single-threaded ufuncs)."""
    import warnings

    idx, n_steps = args
    cfg = _CPU["cfg"]
    fa = dict(cfg["filter_args"], n_steps=n_steps)
    fields = tuple(f[idx] if f.ndim == 3 else f for f in cfg["fields"])
    t0 = time.perf_counter()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if _CPU["kind"] == "reference":  # the unmodified reference package: Filter + _create_filter_func
            from oracle import ref_loader

            ref_loader.ref_filter(cfg["grid_type"], cfg["grid_vars"], fields, **fa)
        else:
            from oracle import np_oracle

            np_oracle.apply_filter(cfg["grid_type"], cfg["grid_vars"], fields, **fa)
    return time.perf_counter() - t0


def cpu_kind():
    """'reference' when the unmodified reference package can be imported here ($GCMF_REFERENCE, /root/reference or
    the pip-installed copy under baseline/_ref), else the numpy oracle port -- and why."""
    try:
        from oracle import ref_loader

        if ref_loader.available():
            ref_loader.load()
            return "reference", f"unmodified reference package at {ref_loader.reference_root()}"
        return "port", "no reference tree ($GCMF_REFERENCE, /root/reference, baseline/_ref): numpy oracle port"
    except Exception as e:  # noqa: BLE001
        return "port", f"reference not importable ({type(e).__name__}: {e}): numpy oracle port"


def cpu_sample(name, nb, n_steps_sample, max_workers=64):
    """Build the bounded CPU sample of a workload: one 2-D slice per host core."""
    cores = len(os.sched_getaffinity(0))
    nbt = workload_nb(name, nb)
    workers = max(1, min(cores, max_workers, nbt or 1))
    cfg = build_workload(name, nb, levels=(0, workers) if nbt else None)
    return cfg, workers, cores


def cpu_baseline(cfg, workers, cores, kind, why, n_steps_sample, pool):
    """Reference path on the host cores: `workers` processes x 1 slice each x n_steps_sample steps."""
    f0 = cfg["fields"][0]
    pts = f0.shape[-1] * f0.shape[-2]
    t0 = time.perf_counter()
    pool.map(_cpu_worker, [(i, n_steps_sample) for i in range(workers)])
    dt = time.perf_counter() - t0
    value = workers * pts * n_steps_sample / dt
    api = ("gcm_filters.Filter + gcm_filters.filter._create_filter_func[_vec] -- the callable xr.apply_ufunc invokes "
           "(xarray is not installed: a 10-line stub stands in for xr.Dataset)") if kind == "reference" else \
        "oracle/np_oracle.py (bit-identical to the reference: tests/test_oracle.py)"
    return {"value": value, "unit": UNIT, "cores": workers, "kind": kind,
            "sample": f"{workers} slice(s) of {f0.shape[-2]}x{f0.shape[-1]} ({cfg['grid_type']}, {f0.dtype}) x "
                      f"n_steps={n_steps_sample} (per-step cost is constant), one process per slice; host has {cores} "
                      f"usable cores; {why}; api: {api}", "seconds": dt}


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


def measured_hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------ GPU arm
class Ctx:
    """Process-group plumbing of one rank."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if args.gpus != self.world and self.world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun for --gpus > 1")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.stream = torch.cuda.current_stream(self.dev)

    def barrier(self):
        self.torch.cuda.synchronize(self.dev)
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def reduce(self, x, op):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=getattr(self.dist.ReduceOp, op))
        return float(t.item())

    def max(self, x):
        return self.reduce(x, "MAX")

    def sum(self, x):
        return self.reduce(x, "SUM")

    def min(self, x):
        return self.reduce(x, "MIN")

    def close(self):
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()


class HostField:
    """ONE host array per direction for the whole field, visible to every rank: a /dev/shm segment that every rank
    maps; a rank pins (cudaHostRegister) and first-touches only its own slab.  Falls back to per-rank pinned slabs
    when /dev/shm cannot hold the field (and says so)."""

    def __init__(self, ctx, tag, shape, dtype, a, b):
        torch = ctx.torch
        self.ctx, self.paths, self.regs, self.maps = ctx, [], [], []
        self.a, self.b = a, b
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self.kind = "per-rank pinned slabs"
        ok = 0.0
        if ctx.world > 1:
            try:
                st = os.statvfs("/dev/shm")
                ok = 1.0 if st.f_bavail * st.f_frsize > 2 * nbytes + (1 << 30) else 0.0
            except OSError:
                ok = 0.0
            ok = ctx.min(ok)
        self.views = {}
        if ok:
            port = os.environ.get("MASTER_PORT", "0")
            for name in ("in", "out"):
                path = f"/dev/shm/gcmf_bench_{port}_{tag}_{name}"
                if ctx.rank == 0:
                    m = np.memmap(path, dtype=dtype, mode="w+", shape=tuple(shape))
                ctx.barrier()
                if ctx.rank != 0:
                    m = np.memmap(path, dtype=dtype, mode="r+", shape=tuple(shape))
                self.paths.append(path)
                self.maps.append(m)
                slab = m[a:b]
                slab[...] = 0  # first touch by the rank that owns the slab
                ptr, n = slab.ctypes.data, slab.nbytes
                rc = torch.cuda.cudart().cudaHostRegister(ptr, n, 0) if n else 0
                if int(rc) != 0:
                    raise RuntimeError(f"cudaHostRegister failed ({rc})")
                if n:
                    self.regs.append(ptr)
                self.views[name] = torch.from_numpy(slab)
            self.kind = "one shared /dev/shm array per direction; each rank pins and first-touches its slab"
        else:
            n = b - a
            for name in ("in", "out"):
                self.views[name] = torch.empty((n,) + tuple(shape[1:]), dtype=getattr(torch, np.dtype(dtype).name),
                                               pin_memory=True)
            if ctx.world > 1:
                self.kind += " (/dev/shm too small for one shared array)"

    def close(self):
        torch = self.ctx.torch
        self.views = {}
        for ptr in self.regs:
            torch.cuda.cudart().cudaHostUnregister(ptr)
        self.regs = []
        self.maps = []
        if self.paths:
            self.ctx.barrier()
            if self.ctx.rank == 0:
                for p in self.paths:
                    try:
                        os.unlink(p)
                    except OSError:
                        pass
        self.paths = []


def host_link_probe(ctx, nbytes=256 << 20, reps=4):
    """Pinned host <-> device copy bandwidth with ALL ranks copying at once (what bounds the end-to-end leg)."""
    torch = ctx.torch
    h_in = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    h_out = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    h_in.fill_(1)
    d_in = torch.empty(nbytes, dtype=torch.uint8, device=ctx.dev)
    d_out = torch.zeros(nbytes, dtype=torch.uint8, device=ctx.dev)
    s1, s2 = torch.cuda.Stream(ctx.dev), torch.cuda.Stream(ctx.dev)
    out = {}
    for mode in ("h2d", "d2h", "duplex"):
        for rep in range(2):  # first pass warms up
            ctx.barrier()
            t0 = time.perf_counter()
            for _ in range(reps):
                if mode in ("h2d", "duplex"):
                    with torch.cuda.stream(s1):
                        d_in.copy_(h_in, non_blocking=True)
                if mode in ("d2h", "duplex"):
                    with torch.cuda.stream(s2):
                        h_out.copy_(d_out, non_blocking=True)
            s1.synchronize()
            s2.synchronize()
            dt = ctx.max(time.perf_counter() - t0)
        moved = reps * nbytes * (2 if mode == "duplex" else 1) * ctx.world
        out[f"{mode}_GBps_all_ranks"] = round(moved / dt / 1e9, 1)
    out["per_gpu_duplex_GBps"] = round(out["duplex_GBps_all_ranks"] / ctx.world, 1)
    out["how"] = f"{nbytes >> 20} MiB pinned buffers x {reps}, all {ctx.world} rank(s) at once, wall clock, max over ranks"
    return out


def make_filter(cfg):
    from gcm_filters_b200 import Filter, FilterShape, GridType

    fa = dict(cfg["filter_args"])
    fa["filter_shape"] = FilterShape[fa["filter_shape"]]
    return Filter(grid_type=GridType[cfg["grid_type"]], grid_vars=cfg["grid_vars"], **fa)


def measure(ctx, args, name, nb=0, primary=False):
    """Device-resident (and, for the primary workload, end-to-end and per-launch) timing of one workload, its batch
    dimension sharded over the ranks.  Returns a dict (rank 0 uses it)."""
    from gcm_filters_b200 import _cabi, engine
    from gcm_filters_b200.filter import _shift_scale
    from gcm_filters_b200.scheduler import batch_slabs

    torch = ctx.torch
    lib = _cabi.get_library()
    nbt = workload_nb(name, nb)
    a, b = batch_slabs(nbt or 1, ctx.world)[ctx.rank]
    slabs = [hi - lo for lo, hi in batch_slabs(nbt or 1, ctx.world)]
    cfg = build_workload(name, nb, levels=(a, b) if nbt else None)
    flt = make_filter(cfg)
    n_steps = int(flt.n_steps)
    lap, spec = flt.laplacian, flt.filter_spec
    c = _shift_scale(spec, lap)
    fields = cfg["fields"]
    ny, nx = fields[0].shape[-2:]
    w = fields[0].dtype.itemsize
    ncomp = len(fields)
    nbl = (b - a) if nbt else (1 if ctx.rank == 0 else 0)  # local batch
    total_nb = nbt or 1
    units = total_nb * ny * nx * n_steps  # grid-point Laplacian steps of one filter call on the WHOLE field
    shape = ((nbt,) if nbt else ()) + (ny, nx)
    res = {"config": config_of(cfg, n_steps, shape, fields[0].dtype), "levels_per_gpu": slabs if nbt else [1],
           "dtype": "f64" if w == 8 else "f32", "n_steps": n_steps}

    dev_in = [torch.from_numpy(np.ascontiguousarray(f)).to(ctx.dev).reshape(max(nbl, 0) if nbt else 1, ny, nx)
              for f in fields] if nbl else []
    dev_out = [torch.empty_like(d) for d in dev_in]
    small = total_nb * ny * nx * w * ncomp <= 2 * 126e6
    flush = torch.empty(int(300e6), dtype=torch.uint8, device=ctx.dev) if small else None

    def run_device():
        if nbl:
            engine.filter_device(lap, spec.p, c, dev_in, dev_out)

    # ---- leg 1: device-resident filter calls ---------------------------------------------------
    for _ in range(args.warmup):
        run_device()
    ctx.barrier()
    sampler = ClockSampler(ctx.local) if (ctx.rank == 0 and primary) else None
    launches0 = lib.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for k in range(args.steps):
        if flush is not None:
            flush.fill_(k & 0xFF)
        ev[k][0].record(ctx.stream)
        run_device()
        ev[k][1].record(ctx.stream)
    ctx.barrier()
    launches = lib.launch_count() - launches0
    res["clocks"] = sampler.stop() if sampler else None
    dev_ms = ctx.max(sum(s.elapsed_time(e) for s, e in ev))
    res["value"] = units * args.steps / (dev_ms * 1e-3)
    res["ms_per_step"] = dev_ms / args.steps
    res["gpu_launches"] = int(ctx.sum(launches))
    b_alg = 5 * w * ncomp + c_bytes_per_point(cfg["grid_type"], w) / max(1, max(slabs) if nbt else 1)
    peak, peak_src = measured_hbm_peak()
    res["whole_call_frac"] = b_alg * res["value"] / ctx.world / 1e9 / peak  # per GPU, whole filter call
    res["algorithmic_bytes_per_pt_step"] = b_alg
    if not primary:
        del dev_in, dev_out
        return res

    # ---- leg 2: end to end through Filter.apply, pinned host in / out --------------------------------
    hf = HostField(ctx, name, (total_nb, ny, nx), fields[0].dtype, a, b) if ncomp == 1 and nbt else None
    if hf is not None:
        host_in, host_out = [hf.views["in"]], [hf.views["out"]]
        if nbl:
            host_in[0].copy_(torch.from_numpy(fields[0]))
    else:
        host_in = [torch.from_numpy(np.ascontiguousarray(f)).pin_memory() for f in fields] if nbl else []
        host_out = [torch.empty_like(h).pin_memory() for h in host_in]

    def e2e_once():
        if not nbl:
            return
        if ncomp == 1:
            flt.apply(host_in[0], dims=["y", "x"], out=host_out[0])
        else:
            flt.apply_to_vector(host_in[0], host_in[1], dims=["y", "x"], out=host_out)

    for _ in range(2):
        e2e_once()
    ctx.barrier()
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_once()  # returns after the D2H copy of the result has completed
    torch.cuda.synchronize(ctx.dev)
    e2e_s = ctx.max(time.perf_counter() - t0)
    if nbl:  # the gathered result equals the device-resident one
        same = bool(torch.equal(torch.nan_to_num(host_out[0].reshape(dev_out[0].shape), nan=-3.0),
                                torch.nan_to_num(dev_out[0].cpu(), nan=-3.0)))
    else:
        same = True
    res["e2e"] = {"value": units * e2e_steps / e2e_s, "unit": UNIT,
                  "h2d_bytes_per_step": int(total_nb * ny * nx * w * ncomp),
                  "d2h_bytes_per_step": int(total_nb * ny * nx * w * ncomp), "steps": e2e_steps,
                  "api": "Filter.apply(pinned host slab, out=pinned host slab) on every rank",
                  "host_array": hf.kind if hf is not None else "per-rank pinned arrays",
                  "matches_device_result": bool(ctx.min(1.0 if same else 0.0) == 1.0)}
    if hf is not None:
        hf.close()
    del host_in, host_out

    # ---- leg 2b: the reference's call signature -- pageable numpy in, numpy out ------------------------
    if not args.no_e2e_numpy:
        np_in = [np.array(f, copy=True) for f in fields] if nbl else []

        def numpy_once():
            if not nbl:
                return None
            if ncomp == 1:
                return flt.apply(np_in[0], dims=["y", "x"])
            return flt.apply_to_vector(np_in[0], np_in[1], dims=["y", "x"])

        numpy_once()
        ctx.barrier()
        reps = 2
        t0 = time.perf_counter()
        for _ in range(reps):
            out_np = numpy_once()
        torch.cuda.synchronize(ctx.dev)
        np_s = ctx.max(time.perf_counter() - t0)
        res["e2e_numpy"] = {"value": units * reps / np_s, "unit": UNIT, "steps": reps,
                            "api": "Filter.apply(numpy array) -> numpy array (pageable memory, pinned staging ring + "
                                   "copy threads inside)"}
        del np_in, out_np
    res["host_link"] = host_link_probe(ctx)

    # ---- leg 3: per-launch durations, same launch sequence as gcmf_filter, CUDA events per launch ----
    if nbl:
        plan = engine.device_plan(lap, ctx.local, fields[0].dtype, ny, nx)
        plan.set_filter(spec.p, c)
        wsb = lib.workspace_bytes(plan.handle, nbl)
        ws = engine.workspace(ctx.dev, wsb)
        kfuse = lib.fused_max_steps(plan.handle) if engine.STEPS_PER_BLOCK != 1 else 0
        nbuf = 4 if kfuse else 2
        bufbytes = wsb // (nbuf * ncomp)
        spec_of = lambda ts: [(t.data_ptr(), nx, ny * nx) for t in ts]  # noqa: E731
        bufs = [[(ws.data_ptr() + (j * ncomp + k) * bufbytes, nx, ny * nx) for k in range(ncomp)] for j in range(nbuf)]
        A, B = bufs[0], bufs[1]
        sptr = ctx.stream.cuda_stream
        records = []
        area = "AREA_WEIGHTED" in cfg["grid_type"]

        def timed(kind, k, fn):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ctx.stream)
            fn()
            e1.record(ctx.stream)
            records.append((kind, k, e0, e1))

        outs = spec_of(dev_out)
        for rep in range(max(1, min(args.steps, 3))):
            X = spec_of(dev_in)
            if area:
                lib.prepare(plan.handle, nbl, X, B, sptr)
                X = B
            if kfuse:  # same blocks as gcmf_filter: ceil(n/kfuse) fused launches, first/last steps included
                cur = 1 if area else 0
                T1, T2 = X, X
                i = 1
                while i <= n_steps:
                    kk = min(n_steps - i + 1, kfuse)
                    O1, O2 = bufs[2 * cur], bufs[2 * cur + 1]
                    kind = "fused_first" if i == 1 else ("fused_last" if i + kk - 1 == n_steps else "fused")
                    timed(kind, kk, lambda: lib.cheb_fused(plan.handle, nbl, i, kk, T1, T2, O1, O2, outs, sptr))
                    T1, T2 = O1, O2
                    cur ^= 1
                    i += kk
                continue
            timed("first", 1, lambda: lib.cheb_step(plan.handle, nbl, 1, X, None, A, outs, sptr))
            T1, T2 = A, X
            for i in range(2, n_steps + 1):
                D = B if (i == 2 and not area) else T2
                timed("mid" if i < n_steps else "last", 1,
                      lambda: lib.cheb_step(plan.handle, nbl, i, T1, T2, D, outs, sptr))
                T2, T1 = T1, D
        torch.cuda.synchronize(ctx.dev)
        per_kind = {}
        for kind, k, e0, e1 in records:
            per_kind.setdefault((kind, k), []).append(e0.elapsed_time(e1))
        dom = max(per_kind, key=lambda key: sum(per_kind[key]))
        dom_ms = float(np.mean(per_kind[dom]))
        b_alg_l = 5 * w * ncomp + c_bytes_per_point(cfg["grid_type"], w) / nbl
        # algorithmic bytes of one launch = B_alg per grid-point step x points x steps the launch performs
        achieved = b_alg_l * nbl * ny * nx * dom[1] / (dom_ms * 1e-3) / 1e9
        fk = (f"cgrid2_kernel ({dom[1]} Chebyshev steps per launch, rows of all arrays streamed through TMA rings)"
              if cfg["grid_type"] == "VECTOR_C_GRID" else
              f"vec2_kernel ({dom[1]} Chebyshev steps per launch, rows streamed through TMA rings)"
              if ncomp == 2 else
              f"march_kernel ({dom[1]} Chebyshev steps per launch, rows streamed through TMA rings, row windows in registers)"
              if (cfg["grid_type"] in ("IRREGULAR_WITH_LAND", "MOM5U", "MOM5T") and w == 8 and nbl >= 2
                  and os.environ.get("GCMF_FUSED_FORM") != "tile") else
              f"fused_kernel ({dom[1]} Chebyshev steps per launch, TMA-staged tiles)")
        kname = {"fused": fk, "fused_first": fk + ", first block", "fused_last": fk + ", last block",
                 "mid": "step_kernel<MODE_MID> (one Chebyshev step)", "first": "step_kernel<MODE_FIRST>",
                 "last": "step_kernel<MODE_LAST>"}[dom[0]]
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": None, "kernel": kname, "ms_per_launch": dom_ms, "steps_per_launch": dom[1],
                    "levels_per_launch": nbl, "algorithmic_bytes_per_pt_step": b_alg_l, "peak_source": peak_src,
                    "launch_mix_ms": {f"{k[0]}x{k[1]}": [len(v), float(np.mean(v))] for k, v in per_kind.items()}}
        prof = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.isfile(prof):  # ncu-measured DRAM bytes per unit of the dominant kernel (committed capture)
            try:
                from gcm_filters_b200 import build as gbuild

                with open(prof) as fh:
                    tr = json.load(fh)
                key = name if dom[0].startswith("fused") else name + "_onestep"
                if key in tr:
                    if tr[key].get("srchash") == gbuild.source_hash(tr[key].get("files")):
                        roofline["traffic"] = tr[key]["bytes_per_pt_step"] * nbl * ny * nx * dom[1]
                        roofline["traffic_source"] = tr[key]["source"]
                    else:
                        roofline["traffic_source"] = ("null: the committed ncu capture (" + tr[key]["source"] +
                                                      ") was taken on a different build of the kernels")
            except Exception:
                pass
        res["roofline"] = roofline
    del dev_in, dev_out
    return res


def free_device(ctx):
    from gcm_filters_b200 import engine

    import gc

    gc.collect()
    engine.release_workspaces()
    engine._pipe_state.clear()
    ctx.torch.cuda.empty_cache()


def banded_secondary(ctx, args, fused=False):
    """cfg5 (VECTOR_C_GRID 2160 x 4320) as latitude bands, one per GPU (strong scaling of one 2-D field).  fused=False:
    one-step kernels that store their border rows straight into the neighbours' peer memory (PeerBandedFilter);
    fused=True: the two-step kernel on bands with two ghost rows per side, which the kernel itself stores into the
    neighbours' peer memory (FusedBandedFilter, exchange="push")."""
    from gcm_filters_b200 import _cabi
    from gcm_filters_b200.scheduler import FusedBandedFilter, PeerBandedFilter

    torch = ctx.torch
    cfg = build_workload("cfg5")
    flt = make_filter(cfg)
    bf = FusedBandedFilter(flt, ctx.rank, ctx.world, exchange="push") if fused else PeerBandedFilter(flt, ctx.rank, ctx.world)
    st = bf.stage(*cfg["fields"])
    lib = _cabi.get_library()
    f0 = cfg["fields"][0]
    ny, nx = f0.shape[-2:]
    steps = max(3, min(args.steps, 10))
    for _ in range(3):
        bf.run(st)
    ctx.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ctx.stream)
    for _ in range(steps):
        bf.run(st)
    e1.record(ctx.stream)
    ctx.barrier()
    ms = ctx.max(e0.elapsed_time(e1))
    st = None
    bf.close()
    n_steps = int(flt.n_steps)
    value = ny * nx * n_steps * steps / (ms * 1e-3)
    peak, _ = measured_hbm_peak()
    b_alg = 5 * 8 * 2 + c_bytes_per_point("VECTOR_C_GRID", 8)
    return {"workload": config_of(cfg, n_steps, (ny, nx), f0.dtype)["workload"], "value": value,
            "ms_per_step": ms / steps, "n_steps": n_steps, "whole_call_frac": b_alg * value / ctx.world / 1e9 / peak,
            "sharding": (f"{ctx.world} latitude band(s); two Chebyshev steps per launch on bands with 2 ghost rows per side, "
                         f"stored by the kernel into the neighbours' peer memory over NVLink as it emits them, "
                         f"flag-synchronised (one launch per two steps, no NCCL on the data path)"
                         if fused else
                         f"{ctx.world} latitude band(s); ghost rows stored by the step kernels into the neighbours' peer "
                         f"memory over NVLink, flag-synchronised (no NCCL on the data path)")}


def gpu_arm(args):
    from gcm_filters_b200 import engine

    ctx = Ctx(args)
    engine.set_steps_per_block(args.steps_per_block)
    cpu = None
    if ctx.rank == 0 and ctx.world == 1 and not args.no_cpu_baseline:
        import multiprocessing as mp

        kind, why = cpu_kind()
        ccfg, workers, cores = cpu_sample(args.workload, args.nb, 8)
        _CPU.update(cfg=ccfg, kind=kind)
        with mp.get_context("fork").Pool(workers) as pool:  # before any CUDA work of this process' children
            cpu = cpu_baseline(ccfg, workers, cores, kind, why, 8, pool)
        _CPU.clear()
        del ccfg
    res = measure(ctx, args, args.workload, args.nb, primary=True)
    secondary = []
    if not args.no_secondary:
        plan = [("cfg2", 0), ("cfg4", 62)] if ctx.world > 1 else [("cfg2", 0), ("cfg4", 0), ("cfg4", 62), ("cfg5", 0),
                                                                  ("cfg1", 0)]
        for name, nb in plan:
            free_device(ctx)
            try:
                r = measure(ctx, args, name, nb)
                secondary.append({"workload": r["config"]["workload"], "value": r["value"], "unit": UNIT,
                                  "ms_per_step": r["ms_per_step"], "n_steps": r["n_steps"], "dtype": r["dtype"],
                                  "whole_call_frac_of_hbm_roofline": r["whole_call_frac"],
                                  "algorithmic_bytes_per_pt_step": r["algorithmic_bytes_per_pt_step"],
                                  "levels_per_gpu": r["levels_per_gpu"]})
            except Exception as e:  # noqa: BLE001  (a secondary line must never cost the primary one)
                secondary.append({"workload": name, "error": f"{type(e).__name__}: {e}"})
        if ctx.world > 1:
            free_device(ctx)
            for fused in (False, True):
                try:
                    secondary.append(banded_secondary(ctx, args, fused=fused))
                except Exception as e:  # noqa: BLE001
                    secondary.append({"workload": "cfg5 banded" + (" (two-step blocks)" if fused else ""),
                                      "error": f"{type(e).__name__}: {e}"})
                free_device(ctx)
    ctx.close()
    if ctx.rank != 0:
        return
    line = {
        "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": ctx.world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": res["dtype"], "data": DATA, "config": res["config"],
        "levels_per_gpu": res["levels_per_gpu"],
        "roofline": res.get("roofline"), "cpu_baseline": cpu, "e2e": res.get("e2e"), "e2e_numpy": res.get("e2e_numpy"),
        "host_link": res.get("host_link"), "gpu_launches": res["gpu_launches"], "clocks": res["clocks"],
        "secondary": secondary,
    }
    emit(line)


# ------------------------------------------------------------------------------------------ banded arm
def banded_arm(args):
    """One 2-D slice split into latitude bands, one band per GPU (BASELINE config 5).  Strong scaling."""
    import torch
    import torch.distributed as dist

    from gcm_filters_b200 import _cabi
    from gcm_filters_b200.scheduler import BandedFilter, FusedBandedFilter, PeerBandedFilter

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = build_workload(args.workload, args.nb)
    flt = make_filter(cfg)
    n_steps = int(flt.n_steps)
    if args.fused:
        bf = FusedBandedFilter(flt, rank, world, exchange="push" if args.push else ("peer" if args.peer else "nccl"))
    else:
        bf = (PeerBandedFilter if args.peer else BandedFilter)(flt, rank, world)
    st = bf.stage(*cfg["fields"])
    f0 = cfg["fields"][0]
    ny, nx = f0.shape[-2:]
    nb = st["nb"]
    H = st.get("H", 1)
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
        "dtype": "f64" if w == 8 else "f32", "data": DATA,
        "config": dict(config_of(cfg, n_steps, f0.shape, f0.dtype),
                       sharding=f"{world} latitude band(s), " +
                       (f"{H} ghost rows, fused {H}-step blocks, ghost rows stored by the kernel into the neighbours' peer memory "
                        f"(NVLink), flag-synchronised" if args.fused and args.push else
                        f"{H} ghost rows, fused {H}-step blocks, ghost rows pulled from peer memory once per block"
                        if args.fused and args.peer else
                        f"{H} ghost rows, fused {H}-step blocks, one NCCL exchange per block" if args.fused else
                        "ghost rows stored by the step kernels into the neighbours' peer memory (NVLink), flag-synchronised"
                        if args.peer else "NCCL send/recv per Chebyshev step")),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "kernel": "step kernels + halo exchange, per GPU (whole filter call)",
                     "algorithmic_bytes_per_pt_step": b_alg, "peak_source": peak_src},
        "cpu_baseline": None, "e2e": None, "gpu_launches": int(launches), "clocks": clocks,
    }
    emit(line)


# ------------------------------------------------------------------------------------------ reference arm
def reference_arm(args):
    """The reference's own CPU implementation of the path on the box's host cores: the UNMODIFIED reference package
    (pip-installed under baseline/_ref in the build container -- it travels to the GPU box -- or $GCMF_REFERENCE /
    /root/reference where present) through Filter + _create_filter_func, else the numpy oracle port with the reason.
    One step = one bounded sample of the workload: one 2-D slice per host core, 6 forced Chebyshev steps."""
    import multiprocessing as mp

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_sample = 6
    kind, why = cpu_kind()
    cfg, workers, cores = cpu_sample(args.workload, args.nb, n_sample)
    from oracle import np_oracle

    fa = cfg["filter_args"]
    n_steps = np_oracle.resolve_n_steps(fa["filter_scale"], fa["dx_min"], fa["filter_shape"])
    nbt = workload_nb(args.workload, args.nb)
    f0 = cfg["fields"][0]
    shape = ((nbt,) if nbt else ()) + tuple(f0.shape[-2:])
    _CPU.update(cfg=cfg, kind=kind)
    t_all, res = 0.0, None
    with mp.get_context("fork").Pool(workers) as pool:
        for _ in range(args.warmup):
            cpu_baseline(cfg, workers, cores, kind, why, 3, pool)
        for _ in range(args.steps):
            res = cpu_baseline(cfg, workers, cores, kind, why, n_sample, pool)
            t_all += res["seconds"]
    units = res["cores"] * f0.shape[-1] * f0.shape[-2] * n_sample
    value = units * args.steps / t_all
    res["value"] = value
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_all / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64" if f0.dtype.itemsize == 8 else "f32",
        "data": DATA, "config": config_of(cfg, n_steps, shape, f0.dtype), "cpu_baseline": res,
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
    ap.add_argument("--nb", type=int, default=0, help="override the batch size (levels / time steps) of the field")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="primary workload only")
    ap.add_argument("--no-e2e-numpy", action="store_true")
    ap.add_argument("--steps-per-block", type=int, default=0, help="0 auto, 1 = one-step kernels only")
    ap.add_argument("--fused", action="store_true",
                    help="with --banded: temporally blocked kernel on the bands, one ghost exchange per 4-step block")
    ap.add_argument("--peer", action="store_true",
                    help="with --banded: ghost rows pushed by the step kernels through peer memory instead of NCCL")
    ap.add_argument("--push", action="store_true",
                    help="with --banded --fused (vector operators): ghost-row exchange fused into the two-step kernel")
    ap.add_argument("--banded", action="store_true",
                    help="latitude-band domain decomposition (strong scaling of one 2-D field; cfg5)")
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
