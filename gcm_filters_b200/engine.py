"""Execution engine: device plans, buffers and streams around the C ABI of libgcmf.so.

PyTorch is used for plumbing only -- device memory (tensors own every pointer handed to the
library), pinned staging, streams.  All arithmetic of the hot path happens inside libgcmf.so.
"""
import os
import threading

import numpy as np

from . import _cabi

_DT = {np.dtype(np.float32): _cabi.GCMF_F32, np.dtype(np.float64): _cabi.GCMF_F64}

# Temporal blocking of the Chebyshev steps: 0 = auto (fuse up to 4 steps per HBM round trip whenever the
# operator/grid has a fused kernel), 1 = never fuse, 2..4 = cap.  Results are bit-identical either way (equal to
# rounding next to a tripolar fold, where mirrored halo cells sum their fluxes in the opposite order).
STEPS_PER_BLOCK = 0


def set_steps_per_block(k):
    global STEPS_PER_BLOCK
    if k not in (0, 1, 2, 3, 4):
        raise ValueError("steps_per_block must be 0 (auto) or 1..4")
    STEPS_PER_BLOCK = int(k)


# In-process multi-GPU scheduling of host-resident batches (the role of dask="parallelized" over batch dims in the
# reference, filter.py:478-486): None = the current device only; a list of device indices = the batch is cut into
# contiguous slabs, one per device, each streamed through its device's copy / filter / copy pipeline by a host thread.
DEVICES = None


def set_devices(devices):
    """``None`` / ``"current"``: filter on the current CUDA device (default).  ``"all"`` or a list of device indices:
    ``Filter.apply`` / ``apply_to_vector`` on HOST arrays (numpy, CPU tensors) shard the flattened batch dimension over
    these devices in-process (one host thread per device, no collective).  Also settable as GCMF_DEVICES=all|0,1,..."""
    global DEVICES
    if devices is None or devices == "current":
        DEVICES = None
        return
    torch = _torch()
    if devices == "all":
        devices = list(range(torch.cuda.device_count()))
    devices = [int(d) for d in devices]
    if not devices or any(d < 0 or d >= torch.cuda.device_count() for d in devices):
        raise ValueError(f"bad device list {devices} ({torch.cuda.device_count()} CUDA device(s) present)")
    DEVICES = devices if len(devices) > 1 else None
    if len(devices) == 1:
        torch.cuda.set_device(devices[0])


def batch_slabs(nb, world):
    """Contiguous slabs of the flattened batch index: sizes differ by at most one (62 levels on 8 GPUs ->
    8,8,8,8,8,8,7,7).  Returns a list of (start, stop)."""
    base, extra = divmod(int(nb), int(world))
    out, start = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append((start, start + n))
        start += n
    return out


def _torch():
    import torch

    if not torch.cuda.is_available():
        raise _cabi.GcmfError("gcm_filters_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch


def _is_torch(a):
    return hasattr(a, "detach") and hasattr(a, "device")


def _float_dtype_of(a):
    """numpy float32 / float64 dtype of an input array; anything else (ints, half, bfloat16) computes in float64."""
    try:
        dt = np.dtype(str(a.dtype).replace("torch.", "")) if _is_torch(a) else np.asarray(a).dtype
    except TypeError:
        return np.dtype(np.float64)
    return dt if dt.kind == "f" and dt.itemsize in (4, 8) else np.dtype(np.float64)


# ------------------------------------------------------------------------------------------
# per-(operator, device, dtype, shape) device state
# ------------------------------------------------------------------------------------------
class DevicePlan:
    """Owns the uploaded coefficient planes and the gcmf_plan handle of one Laplacian on one device."""

    def __init__(self, lap, device_index, np_dtype, ny, nx):
        torch = _torch()
        self.lib = _cabi.get_library()
        self.device = torch.device("cuda", device_index)
        self.np_dtype = np.dtype(np_dtype)
        self.ny, self.nx = ny, nx
        spec = lap._planes
        self.ncomp = lap.ncomp
        self.handle = self.lib.plan_create(spec.op, _DT[self.np_dtype], ny, nx, spec.flags, device_index)
        self.tensors = []  # keep device memory alive
        self.plane_batch_shapes = []
        tdt = torch.float32 if self.np_dtype == np.float32 else torch.float64
        for slot, pl in enumerate(spec.planes):
            if slot == 0 and spec.op == _cabi.OP_REGULAR5:
                pl = spec.mask
                if pl is None:
                    continue
                t = torch.as_tensor(np.require(pl, requirements=["C", "W"]), dtype=torch.uint8).to(self.device)
            else:
                if pl is None:
                    continue
                t = torch.as_tensor(np.require(pl, requirements=["C", "W"])).to(device=self.device, dtype=tdt)
            if tuple(t.shape[-2:]) != (ny, nx):
                raise ValueError(f"grid variable plane has shape {tuple(t.shape)}, field has (..., {ny}, {nx})")
            bshape = tuple(int(s) for s in t.shape[:-2])
            while bshape and bshape[0] == 1:
                bshape = bshape[1:]
            nbp = int(np.prod(bshape)) if bshape else 1
            t = t.reshape((nbp, ny, nx)).contiguous()
            self.tensors.append(t)
            self.plane_batch_shapes.append(bshape)
            self.lib.plan_set_plane(self.handle, slot, t.data_ptr(), nx, ny * nx, nbp)
        self._filter_key = None
        self._spb = 0

    def check_batch(self, batch_shape):
        """Plane batch dims must equal the trailing batch dims of the field (b % plane_nb indexing)."""
        for bs in self.plane_batch_shapes:
            if bs and tuple(batch_shape[len(batch_shape) - len(bs):]) != bs:
                raise ValueError(f"grid variable batch dims {bs} do not match the trailing batch dims of the "
                                 f"field {tuple(batch_shape)}")

    def set_filter(self, p, c):
        if self._spb != STEPS_PER_BLOCK:
            self.lib.set_steps_per_block(self.handle, STEPS_PER_BLOCK)
            self._spb = STEPS_PER_BLOCK
        key = (tuple(float(v) for v in p), float(c))
        if key != self._filter_key:
            self.lib.plan_set_filter(self.handle, key[0], key[1])
            self._filter_key = key

    def __del__(self):
        try:
            self.lib.plan_destroy(self.handle)
        except Exception:
            pass


_ws_lock = threading.Lock()
_workspaces = {}
_launch_locks = {}


def launch_lock(device_index):
    """One launch sequence at a time per device: a filter call is a series of kernel launches that share the
    device workspace and the plan's filter state, so calls from different host threads (e.g. dask's threaded
    scheduler running one block per thread) must not interleave their launches.  Only the (short) host-side
    launch sequence is serialised; the kernels themselves run asynchronously."""
    with _ws_lock:
        lock = _launch_locks.get(device_index)
        if lock is None:
            lock = _launch_locks[device_index] = threading.RLock()
        return lock


def workspace(device, nbytes):
    """Grow-only scratch per device (uint8 tensor, 512-byte aligned by the caching allocator)."""
    torch = _torch()
    with _ws_lock:
        t = _workspaces.get(device.index)
        if t is None or t.numel() < nbytes:
            _workspaces[device.index] = None
            t = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
            _workspaces[device.index] = t
        return t


def fit_batch(plan, nb, device):
    """Largest batch count <= nb whose filter workspace (2 scratch fields per component, 4 with the temporally blocked
    kernels) fits in the device memory that is free right now: a device-resident field close to the HBM size is
    filtered in several passes over batch slices instead of failing to allocate 4x its size."""
    torch = _torch()
    need = plan.lib.workspace_bytes(plan.handle, nb)
    with _ws_lock:
        have = _workspaces.get(device.index)
        have_n = have.numel() if have is not None else 0
    if need <= have_n:
        return nb
    free, _total = torch.cuda.mem_get_info(device)
    free += torch.cuda.memory_reserved(device) - torch.cuda.memory_allocated(device)  # cached blocks are reusable
    budget = int(0.9 * free) + have_n  # the old workspace is dropped before a larger one is allocated
    if need <= budget:
        return nb
    return max(1, min(nb, int(budget // (need / nb))))


def _filter_in_batches(plan, nb, dev_in, dev_out, device, stream):
    """gcmf_filter over the whole batch, or over slices of it when the workspace would not fit (see fit_batch)."""
    step = fit_batch(plan, nb, device)
    ws = workspace(device, plan.lib.workspace_bytes(plan.handle, step))
    for b0 in range(0, nb, step):
        n = min(step, nb - b0)
        plan.lib.filter(plan.handle, n, _specs([t[b0:b0 + n] for t in dev_in]), _specs([t[b0:b0 + n] for t in dev_out]),
                        ws.data_ptr(), ws.numel(), stream)


def release_workspaces():
    with _ws_lock:
        _workspaces.clear()


def device_plan(lap, device_index, np_dtype, ny, nx):
    key = (device_index, np.dtype(np_dtype).str, ny, nx)
    st = lap._device_state.get(key)
    if st is None:
        st = lap._device_state[key] = DevicePlan(lap, device_index, np_dtype, ny, nx)
    return st


# ------------------------------------------------------------------------------------------
# array plumbing
# ------------------------------------------------------------------------------------------
class _Staged:
    """Inputs moved to the device as contiguous (nb, ny, nx) tensors + how to hand results back."""

    def __init__(self, lap, fields):
        torch = _torch()
        if len(fields) != lap.ncomp:
            raise ValueError(f"expected {lap.ncomp} field component(s), got {len(fields)}")
        f0 = fields[0]
        self.kind = "torch" if _is_torch(f0) else "numpy"
        if self.kind == "numpy":
            fields = [np.asarray(getattr(f, "values", f)) for f in fields]
            f0 = fields[0]
        shape = tuple(f0.shape)
        if len(shape) < 2:
            raise ValueError("fields need at least two dimensions (y, x)")
        for f in fields[1:]:
            if tuple(f.shape) != shape:
                raise ValueError("vector components must have the same shape")
        self.shape = shape
        self.ny, self.nx = shape[-2], shape[-1]
        self.batch_shape = shape[:-2]
        self.nb = int(np.prod(self.batch_shape)) if self.batch_shape else 1
        in_dtype = _float_dtype_of(f0)
        if self.kind == "torch":
            self.src_device = f0.device
            self.device = f0.device if f0.device.type == "cuda" else torch.device("cuda", torch.cuda.current_device())
            self.pinned = f0.device.type == "cpu" and f0.is_pinned()
        else:
            self.src_device = None
            self.device = torch.device("cuda", torch.cuda.current_device())
            self.pinned = False
        self.np_dtype = lap.compute_dtype(in_dtype)
        tdt = torch.float32 if self.np_dtype == np.float32 else torch.float64
        self.tdt = tdt
        dev = []
        for f in fields:
            t = f if self.kind == "torch" else torch.from_numpy(np.ascontiguousarray(f))
            t = t.to(device=self.device, dtype=tdt, non_blocking=True)
            dev.append(t.reshape((self.nb, self.ny, self.nx)).contiguous())
        self.dev = dev

    def out_like(self):
        torch = _torch()
        return [torch.empty((self.nb, self.ny, self.nx), dtype=self.tdt, device=self.device) for _ in self.dev]

    def deliver(self, outs, out=None):
        """Hand results back as the caller's kind of array; `out` (tuple of preallocated numpy arrays
        or torch CPU/CUDA tensors, one per component) receives them in place when given."""
        torch = _torch()
        res = []
        if out is not None:
            out = tuple(out) if isinstance(out, (tuple, list)) else (out,)
            for o, dst in zip(outs, out):
                t = dst if _is_torch(dst) else torch.from_numpy(dst)
                t.copy_(o.reshape(self.shape), non_blocking=True)
                res.append(dst)
            torch.cuda.current_stream(self.device).synchronize()
            return tuple(res)
        for o in outs:
            o = o.reshape(self.shape)
            if self.kind == "numpy":
                res.append(o.cpu().numpy())
            elif self.src_device.type == "cpu":
                if self.pinned:
                    h = torch.empty(o.shape, dtype=o.dtype, pin_memory=True)
                    h.copy_(o, non_blocking=True)
                    torch.cuda.current_stream(self.device).synchronize()
                    res.append(h)
                else:
                    res.append(o.cpu())
            else:
                res.append(o)
        return tuple(res)


def _specs(tensors):
    return [(t.data_ptr(), t.shape[-1], t.shape[-2] * t.shape[-1]) for t in tensors]


def run_laplacian(lap, fields):
    """out = Laplacian(fields) on the GPU (one reference ``__call__``)."""
    torch = _torch()
    st = _Staged(lap, fields)
    plan = device_plan(lap, st.device.index, st.np_dtype, st.ny, st.nx)
    plan.check_batch(st.batch_shape)
    outs = st.out_like()
    with torch.cuda.device(st.device):
        stream = torch.cuda.current_stream(st.device).cuda_stream
        plan.lib.laplacian(plan.handle, st.nb, _specs(st.dev), _specs(outs), stream)
    return st.deliver(outs)


# Host-resident inputs are streamed through the device in batch chunks so that the H2D copy of chunk i+1,
# the filter of chunk i and the D2H copy of chunk i-1 overlap (three streams, event-chained).  Copies are
# truly asynchronous only from / to pinned memory (torch CPU tensors with pin_memory, or `out=` pinned).
PIPELINE_MIN_CHUNKS = 4
PIPELINE_TARGET_CHUNKS = int(os.environ.get("GCMF_PIPELINE_CHUNKS", "8"))
PIPELINE_MAX_CHUNK_BYTES = 1 << 30
PIPELINE_NBUF = 2  # device-side input / output chunk buffers in flight
PIPELINE_CACHED_GEOMETRIES = 3
PIPELINE_STAGE_SLOTS = 3   # pinned staging chunks per direction for pageable host arrays
PIPELINE_COPY_THREADS = int(os.environ.get("GCMF_COPY_THREADS", "0")) or max(2, min(8, len(os.sched_getaffinity(0)) // 2))
_pipe_lock = threading.Lock()
_pipe_state = {}


def _host_view(f, nb, ny, nx):
    """(nb, ny, nx) torch CPU view of a host array (numpy or torch), without copying when contiguous."""
    import torch

    if _is_torch(f):
        return f.reshape((nb, ny, nx)) if f.is_contiguous() else f.contiguous().reshape((nb, ny, nx))
    a = np.asarray(getattr(f, "values", f))
    if not a.flags.c_contiguous:
        a = np.ascontiguousarray(a)
    if not a.flags.writeable:
        a = a.copy()
    return torch.from_numpy(a.reshape((nb, ny, nx)))


PIPELINE_LEVEL_GROUP = 2  # the row-streaming kernel marches two batch slices per CTA (gcmf_march.cuh: MARCH_LV)


def _pipeline_chunk(nb, slice_bytes):
    """Batch slices per pipeline chunk: about nb / PIPELINE_TARGET_CHUNKS, capped in bytes, and a multiple of the level
    group of the row-streaming kernel when it is at least one group (measured with three levels per CTA on the cfg3 grid:
    a chunk of 7 slices, 7 of 9 level slots, costs 1.67 ms per slice against 1.33 ms for chunks of 6: e2e 199.5 -> 221.5 G)."""
    chunk = max(1, nb // PIPELINE_TARGET_CHUNKS)
    chunk = min(chunk, max(1, PIPELINE_MAX_CHUNK_BYTES // max(1, slice_bytes)))
    if chunk < PIPELINE_LEVEL_GROUP and nb >= PIPELINE_MIN_CHUNKS * PIPELINE_LEVEL_GROUP:
        chunk = PIPELINE_LEVEL_GROUP  # still at least PIPELINE_MIN_CHUNKS chunks, and every chunk a whole group
    if chunk >= PIPELINE_LEVEL_GROUP:
        whole = PIPELINE_LEVEL_GROUP * (chunk // PIPELINE_LEVEL_GROUP)
        # a whole number of groups: down (more chunks overlap better), except from 3 slices, where down would halve it
        chunk = whole if whole >= 2 * PIPELINE_LEVEL_GROUP or whole == chunk else whole + PIPELINE_LEVEL_GROUP
        while chunk > PIPELINE_LEVEL_GROUP and chunk * slice_bytes > PIPELINE_MAX_CHUNK_BYTES:
            chunk -= PIPELINE_LEVEL_GROUP
    return chunk


def _chunk_schedule(nb, chunk):
    """Batch chunk sizes for the pipeline: small chunks at both ends shorten the ramp (the first H2D copy and
    the last D2H copy cannot overlap with anything), full-size chunks in between keep the launches long."""
    ramp = [s for s in (max(1, chunk // 4), max(1, chunk // 2)) if s < chunk]
    if nb < 2 * sum(ramp) + 2 * chunk:
        ramp = []
    sizes, left = list(ramp), nb - 2 * sum(ramp)
    while left > 0:
        n = min(chunk, left)
        sizes.append(n)
        left -= n
    sizes += ramp[::-1]
    assert sum(sizes) == nb
    return sizes


def _run_filter_pipelined(lap, p, c, fields, out, shape, np_dtype):
    torch = _torch()
    ncomp = lap.ncomp
    ny, nx = shape[-2:]
    nb = int(np.prod(shape[:-2]))
    device = torch.device("cuda", torch.cuda.current_device())
    tdt = torch.float32 if np_dtype == np.float32 else torch.float64
    host_in = [_host_view(f, nb, ny, nx) for f in fields]
    kind_numpy = not _is_torch(fields[0])
    if out is not None:
        out = tuple(out) if isinstance(out, (tuple, list)) else (out,)
        results = list(out)
        host_out = []
        for o in out:  # results must land in the caller's memory: no silent copies
            t = o if _is_torch(o) else torch.from_numpy(o)
            if tuple(t.shape) != tuple(shape) or not t.is_contiguous():
                raise ValueError("`out` must be a C-contiguous array with the shape of the input field")
            host_out.append(t.reshape((nb, ny, nx)))
    else:
        pin = _is_torch(fields[0]) and fields[0].is_pinned()
        host_out = [torch.empty((nb, ny, nx), dtype=tdt, pin_memory=pin) for _ in range(ncomp)]
        results = [h.reshape(shape).numpy() if kind_numpy else h.reshape(shape) for h in host_out]
    for h in host_out:
        if h.dtype != tdt:
            raise ValueError(f"`out` must have dtype {tdt}")
    chunk = _pipeline_chunk(nb, ny * nx * np_dtype.itemsize)
    plan = device_plan(lap, device.index, np_dtype, ny, nx)
    nbuf = PIPELINE_NBUF
    key = (device.index, str(tdt), ncomp, chunk, ny, nx, nbuf)
    with _pipe_lock:
        stt = _pipe_state.pop(key, None)
        if stt is None:
            # a few cached pipeline geometries per process (least recently used goes first): alternating shapes, e.g.
            # two variables filtered in turn from dask worker threads, must not reallocate on every call
            while len(_pipe_state) >= PIPELINE_CACHED_GEOMETRIES:
                _pipe_state.pop(next(iter(_pipe_state)))
            stt = {
                "din": [[torch.empty((chunk, ny, nx), dtype=tdt, device=device) for _ in range(ncomp)] for _ in range(nbuf)],
                "dout": [[torch.empty((chunk, ny, nx), dtype=tdt, device=device) for _ in range(ncomp)] for _ in range(nbuf)],
                "h2d": torch.cuda.Stream(device), "d2h": torch.cuda.Stream(device), "lock": threading.Lock(),
            }
        _pipe_state[key] = stt  # (re-)inserted last = most recently used
    s_comp = torch.cuda.current_stream(device)
    staged_in = not all(h.is_pinned() for h in host_in)
    staged_out = not all(h.is_pinned() for h in host_out)
    with launch_lock(device.index), stt["lock"]:
        plan.set_filter(p, c)
        if staged_in or staged_out:
            return _pipeline_loop_staged(plan, stt, host_in, host_out, results, chunk, nb, ncomp, tdt, device, s_comp,
                                         staged_in, staged_out)
        return _pipeline_loop(plan, stt, host_in, host_out, results, chunk, nb, ncomp, tdt, device, s_comp)


def _pipeline_loop(plan, stt, host_in, host_out, results, chunk, nb, ncomp, tdt, device, s_comp):
    torch = _torch()
    nbuf = len(stt["din"])
    din, dout, s_h2d, s_d2h = stt["din"], stt["dout"], stt["h2d"], stt["d2h"]
    ws = workspace(device, plan.lib.workspace_bytes(plan.handle, chunk))
    ev_h2d = [torch.cuda.Event() for _ in range(nbuf)]
    ev_comp = [torch.cuda.Event() for _ in range(nbuf)]
    ev_d2h = [torch.cuda.Event() for _ in range(nbuf)]
    s_h2d.wait_stream(s_comp)
    b0 = 0
    for i, n in enumerate(_chunk_schedule(nb, chunk)):
        k = i % nbuf
        with torch.cuda.stream(s_h2d):
            if i >= nbuf:
                s_h2d.wait_event(ev_comp[k])  # the filter that read this input buffer has finished
            for cc in range(ncomp):
                src = host_in[cc][b0:b0 + n]
                if src.dtype != tdt:
                    src = src.to(tdt)
                din[k][cc][:n].copy_(src, non_blocking=True)
            ev_h2d[k].record(s_h2d)
        s_comp.wait_event(ev_h2d[k])
        if i >= nbuf:
            s_comp.wait_event(ev_d2h[k])  # the previous result in this output buffer has left the device
        plan.lib.filter(plan.handle, n, _specs([t[:n] for t in din[k]]), _specs([t[:n] for t in dout[k]]),
                        ws.data_ptr(), ws.numel(), s_comp.cuda_stream)
        ev_comp[k].record(s_comp)
        with torch.cuda.stream(s_d2h):
            s_d2h.wait_event(ev_comp[k])
            for cc in range(ncomp):
                host_out[cc][b0:b0 + n].copy_(dout[k][cc][:n], non_blocking=True)
            ev_d2h[k].record(s_d2h)
        b0 += n
    s_d2h.synchronize()
    s_comp.wait_stream(s_d2h)
    return tuple(results)


_copy_pool = None


def _parallel_copy(dst, src):
    """dst[...] = src[...] between two host tensors of the same shape, split over a few threads (numpy releases the
    GIL while it copies): one thread moves ~10 GB/s, a pageable 4 GB field each way would otherwise dominate."""
    global _copy_pool
    d = dst.numpy().reshape(-1)
    s_ = src.numpy() if _is_torch(src) else src
    s_ = s_.reshape(-1)
    n = d.shape[0]
    parts = PIPELINE_COPY_THREADS if n >= (1 << 20) else 1
    if parts == 1:
        np.copyto(d, s_, casting="unsafe")
        return
    if _copy_pool is None:
        from concurrent.futures import ThreadPoolExecutor

        _copy_pool = ThreadPoolExecutor(max_workers=PIPELINE_COPY_THREADS, thread_name_prefix="gcmf-copy")
    step = (n + parts - 1) // parts
    futs = [_copy_pool.submit(np.copyto, d[a:a + step], s_[a:a + step], "unsafe") for a in range(0, n, step)]
    for f in futs:
        f.result()


def _pipeline_loop_staged(plan, stt, host_in, host_out, results, chunk, nb, ncomp, tdt, device, s_comp, staged_in,
                          staged_out):
    """The chunk pipeline for PAGEABLE host arrays (plain numpy in / numpy out: what an xarray user passes).
    cudaMemcpyAsync from / to pageable memory is synchronous and staged by the driver through one small buffer;
    here a producer thread copies chunk i+1 of the input into a ring of pinned staging chunks while chunk i is in
    flight, and a consumer thread copies finished chunks out of a pinned ring into the caller's array, so that H2D,
    filter and D2H overlap exactly as they do for pinned arrays."""
    import queue

    torch = _torch()
    nbuf = len(stt["din"])
    din, dout, s_h2d, s_d2h = stt["din"], stt["dout"], stt["h2d"], stt["d2h"]
    ws = workspace(device, plan.lib.workspace_bytes(plan.handle, chunk))
    ny, nx = din[0][0].shape[-2:]
    slots = PIPELINE_STAGE_SLOTS
    skey = ("stage", slots, staged_in, staged_out)
    if skey not in stt:
        def ring():
            return [[torch.empty((chunk, ny, nx), dtype=tdt, pin_memory=True) for _ in range(ncomp)] for _ in range(slots)]
        stt[skey] = (ring() if staged_in else None, ring() if staged_out else None)
    ring_in, ring_out = stt[skey]
    sizes = _chunk_schedule(nb, chunk)
    offs = [0]
    for n in sizes:
        offs.append(offs[-1] + n)
    nchunks = len(sizes)
    ev_h2d = [torch.cuda.Event() for _ in range(nchunks)]
    ev_comp = [torch.cuda.Event() for _ in range(nbuf)]
    ev_d2h = [torch.cuda.Event() for _ in range(nchunks)]
    h2d_recorded = [threading.Event() for _ in range(nchunks)]
    q_in, q_out = queue.Queue(), queue.Queue()
    free_out = threading.Semaphore(slots)
    errors = []

    def producer():
        try:
            for i, n in enumerate(sizes):
                if i >= slots:  # the H2D copy that last read this staging slot has finished
                    h2d_recorded[i - slots].wait()
                    ev_h2d[i - slots].synchronize()
                for cc in range(ncomp):
                    _parallel_copy(ring_in[i % slots][cc][:n], host_in[cc][offs[i]:offs[i] + n])
                q_in.put(i)
        except BaseException as e:  # noqa: BLE001
            errors.append(e)
            q_in.put(-1)

    def consumer():
        try:
            while True:
                i = q_out.get()
                if i is None:
                    return
                ev_d2h[i].synchronize()
                n = sizes[i]
                for cc in range(ncomp):
                    _parallel_copy(host_out[cc][offs[i]:offs[i] + n], ring_out[i % slots][cc][:n])
                free_out.release()
        except BaseException as e:  # noqa: BLE001
            errors.append(e)
            free_out.release()

    threads = []
    if staged_in:
        threads.append(threading.Thread(target=producer, name="gcmf-stage-in", daemon=True))
    if staged_out:
        threads.append(threading.Thread(target=consumer, name="gcmf-stage-out", daemon=True))
    for t in threads:
        t.start()
    s_h2d.wait_stream(s_comp)
    try:
        for i, n in enumerate(sizes):
            k = i % nbuf
            b0 = offs[i]
            if staged_in:
                got = q_in.get()
                if got < 0:
                    raise errors[0]
            with torch.cuda.stream(s_h2d):
                if i >= nbuf:
                    s_h2d.wait_event(ev_comp[k])  # the filter that read this input buffer has finished
                for cc in range(ncomp):
                    src = ring_in[i % slots][cc][:n] if staged_in else host_in[cc][b0:b0 + n]
                    din[k][cc][:n].copy_(src, non_blocking=True)
                ev_h2d[i].record(s_h2d)
            h2d_recorded[i].set()
            s_comp.wait_event(ev_h2d[i])
            if i >= nbuf:
                s_comp.wait_event(ev_d2h[i - nbuf])  # the previous result in this output buffer has left the device
            plan.lib.filter(plan.handle, n, _specs([t[:n] for t in din[k]]), _specs([t[:n] for t in dout[k]]),
                            ws.data_ptr(), ws.numel(), s_comp.cuda_stream)
            ev_comp[k].record(s_comp)
            if staged_out:
                free_out.acquire()  # a pinned staging chunk is free again (back-pressure from the consumer)
                if errors:
                    raise errors[0]
            with torch.cuda.stream(s_d2h):
                s_d2h.wait_event(ev_comp[k])
                for cc in range(ncomp):
                    dst = ring_out[i % slots][cc][:n] if staged_out else host_out[cc][b0:b0 + n]
                    dst.copy_(dout[k][cc][:n], non_blocking=True)
                ev_d2h[i].record(s_d2h)
            if staged_out:
                q_out.put(i)
    finally:
        for ev in h2d_recorded:
            ev.set()
        if staged_out:
            q_out.put(None)
        for t in threads:
            t.join()
    s_d2h.synchronize()
    s_comp.wait_stream(s_d2h)
    if errors:
        raise errors[0]
    return tuple(results)


def _wants_pipeline(lap, fields, out):
    """Stream the batch through the device in chunks?  Only for host-resident inputs (and outputs) with a
    batch axis long enough to overlap copies with compute, and 2-D (shared) coefficient planes."""
    f0 = fields[0]
    if _is_torch(f0) and f0.device.type != "cpu":
        return None
    if out is not None:
        o0 = out[0] if isinstance(out, (tuple, list)) else out
        if _is_torch(o0) and o0.is_cuda:
            return None
    shape = tuple(f0.shape)
    if len(shape) <= 2 or len(fields) != lap.ncomp or any(tuple(f.shape) != shape for f in fields):
        return None
    np_dtype = lap.compute_dtype(_float_dtype_of(f0))
    planes = [pl for pl in lap._planes.planes if pl is not None] + ([lap._planes.mask] if lap._planes.mask is not None else [])
    if any(np.ndim(pl) > 2 for pl in planes):
        return None
    nb = int(np.prod(shape[:-2]))
    if nb < PIPELINE_MIN_CHUNKS * _pipeline_chunk(nb, shape[-2] * shape[-1] * np_dtype.itemsize):
        return None
    return shape, np_dtype


def run_area_op(lap, field, divide):
    """AreaWeightedMixin.prepare (field * area) / finalize (field / area) on the GPU (kernels.py:100-104)."""
    torch = _torch()
    st = _Staged(lap, (field,))
    plan = device_plan(lap, st.device.index, st.np_dtype, st.ny, st.nx)
    plan.check_batch(st.batch_shape)
    outs = st.out_like()
    with launch_lock(st.device.index), torch.cuda.device(st.device):
        stream = torch.cuda.current_stream(st.device).cuda_stream
        op = plan.lib.finalize if divide else plan.lib.prepare
        op(plan.handle, st.nb, _specs(st.dev), _specs(outs), stream)
    return st.deliver(outs)[0]


def _devices_from_env():
    global DEVICES
    env = os.environ.get("GCMF_DEVICES")
    if env and DEVICES is None:
        set_devices("all" if env.strip() == "all" else [int(v) for v in env.split(",") if v.strip()])
        os.environ.pop("GCMF_DEVICES", None)


def _run_filter_multi_device(lap, p, c, fields, out, devices):
    """Host-resident batch sharded over several devices of this process: slab r of the flattened batch goes through
    device r's pipeline from its own host thread; results land in one output array.  No collective."""
    torch = _torch()
    f0 = fields[0]
    shape = tuple(f0.shape)
    ny, nx = shape[-2:]
    nb = int(np.prod(shape[:-2]))
    np_dtype = lap.compute_dtype(_float_dtype_of(f0))
    tdt = torch.float32 if np_dtype == np.float32 else torch.float64
    host_in = [_host_view(f, nb, ny, nx) for f in fields]
    kind_numpy = not _is_torch(f0)
    if out is not None:
        outs = tuple(out) if isinstance(out, (tuple, list)) else (out,)
        results = list(outs)
        host_out = []
        for o in outs:
            t = o if _is_torch(o) else torch.from_numpy(o)
            if tuple(t.shape) != shape or not t.is_contiguous() or t.dtype != tdt:
                raise ValueError(f"`out` must be a C-contiguous {tdt} array with the shape of the input field")
            host_out.append(t.reshape((nb, ny, nx)))
    else:
        pin = _is_torch(f0) and f0.is_pinned()
        host_out = [torch.empty((nb, ny, nx), dtype=tdt, pin_memory=pin) for _ in range(lap.ncomp)]
        results = [h.reshape(shape).numpy() if kind_numpy else h.reshape(shape) for h in host_out]
    slabs = [(d, a, b) for d, (a, b) in zip(devices, batch_slabs(nb, len(devices))) if b > a]
    errors = []

    def work(d, a, b):
        try:
            torch.cuda.set_device(d)
            run_filter(lap, p, c, tuple(h[a:b] for h in host_in), out=tuple(h[a:b] for h in host_out), _devices=())
        except BaseException as e:  # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=work, args=sl, name=f"gcmf-dev{sl[0]}", daemon=True) for sl in slabs]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return tuple(results)


def run_filter(lap, p, c, fields, out=None, _devices=None):
    """filtered = filter_func(fields) on the GPU: prepare, n_steps Chebyshev steps, finalize."""
    torch = _torch()
    _devices_from_env()
    devices = DEVICES if _devices is None else _devices
    f0 = fields[0]
    if devices and len(devices) > 1 and len(f0.shape) > 2 and not (_is_torch(f0) and f0.device.type != "cpu") and \
            len(fields) == lap.ncomp and int(np.prod(f0.shape[:-2])) >= 2 and \
            not any(np.ndim(pl) > 2 for pl in list(lap._planes.planes) + [lap._planes.mask] if pl is not None):
        return _run_filter_multi_device(lap, p, c, fields, out, devices)
    piped = _wants_pipeline(lap, fields, out)
    if piped is not None:
        return _run_filter_pipelined(lap, p, c, fields, out, *piped)
    st = _Staged(lap, fields)
    plan = device_plan(lap, st.device.index, st.np_dtype, st.ny, st.nx)
    plan.check_batch(st.batch_shape)
    outs = st.out_like()
    with launch_lock(st.device.index), torch.cuda.device(st.device):
        plan.set_filter(p, c)
        stream = torch.cuda.current_stream(st.device).cuda_stream
        _filter_in_batches(plan, st.nb, st.dev, outs, st.device, stream)
    return st.deliver(outs, out)


def filter_device(lap, p, c, dev_in, dev_out):
    """Device-resident entry used by the scheduler and the benchmark: (nb, ny, nx) CUDA tensors in,
    preallocated CUDA tensors out, nothing but kernel launches on the current stream."""
    torch = _torch()
    t0 = dev_in[0]
    nb, ny, nx = (int(s) for s in t0.shape)
    np_dtype = np.dtype(np.float32 if t0.dtype == torch.float32 else np.float64)
    plan = device_plan(lap, t0.device.index, np_dtype, ny, nx)
    with launch_lock(t0.device.index):
        plan.set_filter(p, c)
        stream = torch.cuda.current_stream(t0.device).cuda_stream
        _filter_in_batches(plan, nb, dev_in, dev_out, t0.device, stream)
    return dev_out
