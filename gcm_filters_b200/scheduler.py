"""Multi-GPU scheduling of the filter path: one process per GPU (torchrun), ``torch.distributed`` for plumbing.

Replaces the role of ``xr.apply_ufunc(..., dask="parallelized")`` (reference ``filter.py:478-486``), which
can only chunk batch dimensions and never splits the two filtered dimensions.

* :func:`batch_slabs` / :func:`apply_batch_sharded` -- the natural sharding: every 2-D slice is filtered
  independently, so the flattened batch index (time x depth ...) is cut into contiguous slabs, one per
  rank; coefficient planes are replicated; **no data-path collective** (an optional all-gather hands
  every rank the full result).
* :class:`BandedFilter` -- latitude-band domain decomposition for a slice that should not (or cannot)
  live on one device (BASELINE config 5): rank r owns rows [j0, j1) plus one ghost row on each side and
  exchanges the ghost rows of ``T_{i-1}`` with its two neighbours after every Chebyshev step through
  NCCL point-to-point (``batch_isend_irecv``) over NVLink.  ``T_{i-2}`` and the running ``bar`` are
  point-wise and need no halo.  y is periodic for the non-tripolar grids, so the bands form a ring; for
  tripolar grids the top band folds onto itself locally and row 0 is land, so there is no wrap link.
* :class:`PeerBandedFilter` -- the same decomposition with the exchange fused into the step kernels: band
  buffers in symmetric memory, border rows stored straight into the neighbours' ghost rows over NVLink,
  flag-synchronised (``gcmf_cheb_step_halo``).  One launch per step and rank, no NCCL on the data path.
* :class:`FusedBandedFilter` -- bands driven by the temporally blocked kernel (scalar operators, periodic
  grids): four ghost rows, one exchange per 4-step block, by NCCL or by pulling from peer memory.

Measured strong scaling: cfg5 on 8 GPUs 2.0x with NCCL per step vs 5.9x with the fused
peer-memory exchange; cfg3 (8 levels) with fused bands 5.7x on 8 GPUs.
"""
import numpy as np

from . import _cabi
from . import engine
from .filter import _shift_scale

_AREA_FLAG = _cabi.FLAG_AREA


batch_slabs = engine.batch_slabs  # contiguous slabs of the flattened batch index, sizes differ by at most one


def band_rows(ny, world):
    """Contiguous latitude bands (row ranges) of a ny-row grid."""
    return batch_slabs(ny, world)


def apply_batch_sharded(apply_fn, field, rank, world, gather=False, group=None):
    """Filter this rank's slab of ``field`` (shape (..., ny, nx); leading axes are batch) with
    ``apply_fn`` (e.g. ``lambda a: flt.apply(a, dims)``).  Returns the filtered slab, shaped
    (slab, ny, nx), or -- with ``gather=True`` -- the full result on every rank (all-gather of slabs
    padded to the largest slab; the only collective, and not on the data path of the filter)."""
    shape = tuple(field.shape)
    ny, nx = shape[-2:]
    flat = field.reshape((-1, ny, nx))
    nb = flat.shape[0]
    a, b = batch_slabs(nb, world)[rank]
    local = apply_fn(flat[a:b]) if b > a else flat[a:b]
    if not gather:
        return local
    import torch
    import torch.distributed as dist

    was_numpy = not engine._is_torch(local)
    t = torch.as_tensor(local)
    if dist.get_backend(group) == "nccl" and not t.is_cuda:  # NCCL moves device memory only
        t = t.cuda()
    smax = max(hi - lo for lo, hi in batch_slabs(nb, world))
    pad = torch.zeros((smax, ny, nx), dtype=t.dtype, device=t.device)
    pad[: b - a] = t
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    full = torch.cat([parts[r][: hi - lo] for r, (lo, hi) in enumerate(batch_slabs(nb, world))]).reshape(shape)
    return full.cpu().numpy() if was_numpy else full


class BandedFilter:
    """One rank's share of a latitude-band decomposed filter.

    ``library`` / ``device`` default to the CUDA build of libgcmf.so and the current CUDA device.  The
    CPU test-suite passes the host emulator of the same C ABI and ``device="cpu"`` to exercise the band
    bookkeeping and the halo exchange over gloo; the product never does.
    """

    def __init__(self, flt, rank, world, group=None, library=None, device=None):
        import torch

        self.flt, self.rank, self.world, self.group = flt, int(rank), int(world), group
        self.lap = flt.laplacian
        self.lib = library if library is not None else _cabi.get_library()
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        self.spec = flt.filter_spec
        self.c = _shift_scale(self.spec, self.lap)
        self._plans = {}

    # ---- band-local plan: planes sliced with one ghost row above and below ---------------------
    def _plan(self, np_dtype, ny, nx, halo=1):
        import torch

        key = (np.dtype(np_dtype).str, ny, nx, halo)
        if key in self._plans:
            return self._plans[key]
        j0, j1 = band_rows(ny, self.world)[self.rank]
        nyl = j1 - j0
        spec = self.lap._planes
        flags = spec.flags & ~_cabi.FLAG_WRAP_Y  # ghost rows are real memory
        if self.rank != self.world - 1:
            flags &= ~_cabi.FLAG_FOLD_N
        if self.rank != 0:
            flags &= ~_cabi.FLAG_CUT_S
        dt = engine._DT[np.dtype(np_dtype)]
        dev_index = self.device.index if self.device.type == "cuda" else 0
        h = self.lib.plan_create(spec.op, dt, nyl, nx, flags, dev_index)
        rows = np.arange(j0 - halo, j1 + halo) % ny
        tdt = torch.float32 if np.dtype(np_dtype) == np.float32 else torch.float64
        keep = []
        for slot, pl in enumerate(spec.planes):
            is_mask = slot == 0 and spec.op == _cabi.OP_REGULAR5
            src = spec.mask if is_mask else pl
            if src is None:
                continue
            src = np.asarray(src)
            if src.ndim != 2:
                raise NotImplementedError("band decomposition supports 2-D grid variables")
            band = np.ascontiguousarray(src[rows])  # (nyl + 2*halo, nx), periodic ghost rows
            t = torch.as_tensor(band).to(device=self.device, dtype=torch.uint8 if is_mask else tdt).contiguous()
            keep.append(t)
            self.lib.plan_set_plane(h, slot, t.data_ptr() + halo * nx * t.element_size(), nx, (nyl + 2 * halo) * nx, 1)
        self.lib.plan_set_filter(h, [float(v) for v in self.spec.p], float(self.c))
        self._plans[key] = (h, keep, j0, j1, flags)
        return self._plans[key]

    # ---- ghost-row exchange of a (ncomp, nb, nyl+2h, nx) tensor with h ghost rows per side ---------------
    def _exchange(self, t, ring, h=1):
        import torch
        import torch.distributed as dist

        nyl = t.shape[-2] - 2 * h
        north, south = (self.rank + 1) % self.world, (self.rank - 1) % self.world
        has_n = ring or self.rank != self.world - 1
        has_s = ring or self.rank != 0
        if self.world == 1:
            if ring:
                t[..., 0:h, :] = t[..., nyl:nyl + h, :]
                t[..., nyl + h:nyl + 2 * h, :] = t[..., h:2 * h, :]
            return
        top = t[..., nyl:nyl + h, :].contiguous()  # my northernmost owned rows -> north neighbour's south ghosts
        bot = t[..., h:2 * h, :].contiguous()      # my southernmost owned rows -> south neighbour's north ghosts
        gn, gs = torch.empty_like(top), torch.empty_like(bot)
        ops = []
        # Ordering matters when north == south (world == 2): NCCL matches the k-th send to a peer with the
        # k-th receive from it, so sends go (bottom rows, top rows) against receives (north ghosts, south
        # ghosts).  gloo matches by tag.
        if has_s:
            ops.append(dist.P2POp(dist.isend, bot, south, group=self.group, tag=2))
        if has_n:
            ops.append(dist.P2POp(dist.irecv, gn, north, group=self.group, tag=2))
            ops.append(dist.P2POp(dist.isend, top, north, group=self.group, tag=1))
        if has_s:
            ops.append(dist.P2POp(dist.irecv, gs, south, group=self.group, tag=1))
        for req in dist.batch_isend_irecv(ops) if ops else []:
            req.wait()
        if has_n:
            t[..., nyl + h:nyl + 2 * h, :] = gn
        if has_s:
            t[..., 0:h, :] = gs

    # ---- the filter on this rank's band -----------------------------------------------------------
    def stage(self, *fields):
        """Move this rank's band of the GLOBAL component arrays (numpy, shape (..., ny, nx)) to the device.
        Returns the state :meth:`run` works on."""
        import torch

        lap = self.lap
        ncomp = lap.ncomp
        assert len(fields) == ncomp
        f0 = np.asarray(fields[0])
        ny, nx = f0.shape[-2:]
        np_dtype = lap.compute_dtype(f0.dtype if f0.dtype.kind == "f" else np.float64)
        h, _keep, j0, j1, flags = self._plan(np_dtype, ny, nx)
        nyl = j1 - j0
        nb = int(np.prod(f0.shape[:-2])) if f0.ndim > 2 else 1
        tdt = torch.float32 if np_dtype == np.float32 else torch.float64

        def new():
            return torch.zeros((ncomp, nb, nyl + 2, nx), dtype=tdt, device=self.device)

        X0 = new()
        for k, f in enumerate(fields):
            band = np.ascontiguousarray(np.asarray(f).reshape((nb, ny, nx))[:, j0:j1])
            X0[k, :, 1:nyl + 1] = torch.as_tensor(band).to(device=self.device, dtype=tdt)
        return dict(h=h, flags=flags, j0=j0, j1=j1, nyl=nyl, nb=nb, nx=nx, ncomp=ncomp, X0=X0, X=new(), A=new(), B=new(),
                    bar=torch.empty((ncomp, nb, nyl, nx), dtype=tdt, device=self.device), batch_shape=f0.shape[:-2])

    def run(self, st):
        """All n_steps Chebyshev steps on the device-resident band, one ghost-row exchange per step.
        Nothing but kernel launches, small pack/unpack copies and NCCL point-to-point calls."""
        import torch

        lap, lib = self.lap, self.lib
        h, flags, nyl, nb, nx, ncomp = st["h"], st["flags"], st["nyl"], st["nb"], st["nx"], st["ncomp"]
        ring = bool(lap._planes.flags & _cabi.FLAG_WRAP_Y) and not (lap._planes.flags & _cabi.FLAG_CUT_S)
        n = int(self.spec.n_steps)
        X, A, B, bar = st["X"], st["A"], st["B"], st["bar"]
        X.copy_(st["X0"])
        es = X.element_size()

        def inner(t):  # gcmf_field specs addressing the owned rows (row 1 of the ghosted array)
            return [(t[k].data_ptr() + nx * es, nx, (nyl + 2) * nx) for k in range(ncomp)]

        def plain(t):
            return [(t[k].data_ptr(), nx, nyl * nx) for k in range(ncomp)]

        stream = torch.cuda.current_stream(self.device).cuda_stream if self.device.type == "cuda" else 0
        if flags & _AREA_FLAG:
            lib.prepare(h, nb, inner(X), inner(B), stream)
            X, B = B, X
        self._exchange(X, ring)
        lib.cheb_step(h, nb, 1, inner(X), None, inner(A), plain(bar), stream)
        self._exchange(A, ring)
        T1, T2 = A, X
        for i in range(2, n + 1):
            D = T2  # all band buffers are ours: T_i overwrites T_{i-2} in place (pointer rotation)
            lib.cheb_step(h, nb, i, inner(T1), inner(T2), inner(D), plain(bar), stream)
            if i < n:
                self._exchange(D, ring)
            T2, T1 = T1, D
        return bar

    def apply(self, *fields):
        """``fields``: the GLOBAL component arrays (numpy, shape (..., ny, nx)); every rank passes the same
        arrays and keeps only its band (+ ghost rows).  Returns this rank's rows [j0, j1) of the filtered
        component(s) as numpy arrays, and (j0, j1)."""
        import torch

        st = self.stage(*fields)
        bar = self.run(st)
        if self.device.type == "cuda":
            torch.cuda.current_stream(self.device).synchronize()
        outs = tuple(bar[k].reshape(tuple(st["batch_shape"]) + (st["nyl"], st["nx"])).cpu().numpy()
                     for k in range(st["ncomp"]))
        return outs, (st["j0"], st["j1"])


class PeerBandedFilter(BandedFilter):
    """Latitude-band decomposition whose ghost-row exchange is fused into the step kernels.

    The band buffers live in *symmetric memory* (``torch.distributed._symmetric_memory``): every rank maps
    its neighbours' buffers, so ``gcmf_cheb_step_halo`` can store the first / last owned row of each new
    ``T_i`` straight into the neighbouring GPU's ghost row over NVLink and raise a flag there; the next
    step on that GPU spins on the flag only in the CTAs that touch ghost rows.  Per Chebyshev step this is
    ONE kernel launch per rank and no NCCL call (the :class:`BandedFilter` base class needs two pack
    kernels, a grouped send/recv and two unpack kernels).

    ``memory="local"`` (the default for world == 1 and for the host-emulator tests) keeps the buffers in
    ordinary memory: a single rank is its own north and south neighbour.
    """

    FLAG_WORDS = 16  # uint32: [0] from_south, [1] from_north, [2..3] counters, rest padding

    def __init__(self, flt, rank, world, group=None, library=None, device=None, memory=None):
        super().__init__(flt, rank, world, group=group, library=library, device=device)
        self.memory = memory or ("symmetric" if self.world > 1 else "local")
        self._epoch = 0
        self._keep = []
        self._buffers = {}

    def close(self):
        """Drop the symmetric-memory mappings.  Call before ``destroy_process_group()``: the handles must not
        outlive the process group they were established on."""
        import gc

        self._buffers.clear()
        self._keep.clear()
        gc.collect()

    def _alloc(self, nbytes):
        """(local uint8 tensor, [base pointer of that buffer on every rank])"""
        import torch

        if nbytes in self._buffers:  # one allocation (and one rendezvous) per geometry
            t, ptrs = self._buffers[nbytes]
            if self.memory != "local":
                import torch.distributed as dist

                # Order matters: a rank's last step only waits for its neighbours' previous step, so a neighbour's
                # final kernel may still be signalling into my flag words when I get here.  Everybody first finishes
                # its own kernels and meets at the barrier -- nobody is writing into anybody's buffer any more -- and
                # only then are the flags and ghost rows cleared and the epoch reset; a second barrier keeps a fast
                # rank from pushing into a buffer that a slow rank has not cleared yet.
                torch.cuda.synchronize(self.device)
                dist.barrier(group=self.group)
                t.zero_()
                self._epoch = 0
                torch.cuda.synchronize(self.device)
                dist.barrier(group=self.group)
            else:
                t.zero_()
                self._epoch = 0
            return t, ptrs
        if self.memory == "local":
            assert self.world == 1, "local memory only supports a single rank (its own neighbour)"
            t = torch.zeros(nbytes, dtype=torch.uint8, device=self.device)
            self._buffers[nbytes] = (t, [t.data_ptr()])
            return self._buffers[nbytes]
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem

        t = symm_mem.empty(nbytes, dtype=torch.uint8, device=self.device)
        t.zero_()
        hdl = symm_mem.rendezvous(t, self.group if self.group is not None else dist.group.WORLD)
        self._keep.append(hdl)
        self._buffers[nbytes] = (t, [int(p) for p in hdl.buffer_ptrs])
        return self._buffers[nbytes]

    def stage(self, *fields):
        import torch

        lap = self.lap
        ncomp = lap.ncomp
        assert len(fields) == ncomp
        f0 = np.asarray(fields[0])
        ny, nx = f0.shape[-2:]
        np_dtype = lap.compute_dtype(f0.dtype if f0.dtype.kind == "f" else np.float64)
        h, _keep, j0, j1, flags = self._plan(np_dtype, ny, nx)
        nyl = j1 - j0
        nb = int(np.prod(f0.shape[:-2])) if f0.ndim > 2 else 1
        tdt = torch.float32 if np_dtype == np.float32 else torch.float64
        es = np_dtype.itemsize
        bands = band_rows(ny, self.world)
        nyl_max = max(b - a for a, b in bands)
        rows = nyl_max + 2                      # common allocation: ghost row, owned rows, ghost row
        slab = ncomp * nb * rows * nx           # elements of one ghosted array
        slab_bytes = (slab * es + 255) // 256 * 256
        flag_off = 3 * slab_bytes
        total = flag_off + 4 * self.FLAG_WORDS
        buf, bases = self._alloc(total)
        if self.device.type == "cuda":
            torch.cuda.synchronize(self.device)

        def view(k):
            return buf[k * slab_bytes:k * slab_bytes + slab * es].view(tdt).view(ncomp, nb, rows, nx)

        X, A, B = view(0), view(1), view(2)
        X0 = torch.zeros((ncomp, nb, nyl, nx), dtype=tdt, device=self.device)
        for k, f in enumerate(fields):
            band = np.ascontiguousarray(np.asarray(f).reshape((nb, ny, nx))[:, j0:j1])
            X0[k] = torch.as_tensor(band).to(device=self.device, dtype=tdt)
        ring = bool(lap._planes.flags & _cabi.FLAG_WRAP_Y) and not (lap._planes.flags & _cabi.FLAG_CUT_S)
        north = (self.rank + 1) % self.world if (ring or self.rank != self.world - 1) else None
        south = (self.rank - 1) % self.world if (ring or self.rank != 0) else None
        st = dict(h=h, flags=flags, j0=j0, j1=j1, nyl=nyl, nb=nb, nx=nx, ncomp=ncomp, X0=X0, X=X, A=A, B=B, buf=buf,
                  bar=torch.empty((ncomp, nb, nyl, nx), dtype=tdt, device=self.device), batch_shape=f0.shape[:-2],
                  bases=bases, slab_bytes=slab_bytes, flag_off=flag_off, rows=rows, es=es, north=north, south=south,
                  nyl_north=(bands[north][1] - bands[north][0]) if north is not None else 0,
                  nyl_south=(bands[south][1] - bands[south][0]) if south is not None else 0)
        return st

    def _halo(self, st, which, wait_value, signal_value, push=True):
        """gcmf_halo for the array in slab `which` (0 = X, 1 = A, 2 = B)."""
        es, nx, rows, nb, ncomp = st["es"], st["nx"], st["rows"], st["nb"], st["ncomp"]
        comp = nb * rows * nx * es  # bytes of one component of a ghosted array
        me = st["bases"][self.rank if self.memory != "local" else 0]
        hl = _cabi.Halo()
        hl.north_bstride = hl.south_bstride = rows * nx
        fl = me + st["flag_off"]
        hl.counters = fl + 8
        hl.wait_value, hl.signal_value = wait_value, signal_value
        for side, peer in (("north", st["north"]), ("south", st["south"])):
            if peer is None:
                continue
            pb = st["bases"][peer if self.memory != "local" else 0]
            arr = pb + which * st["slab_bytes"]
            if side == "north":  # my row ny-1 -> north neighbour's row -1 (index 0 of its ghosted array)
                if push:
                    for k in range(ncomp):
                        hl.north_ghost[k] = arr + k * comp
                hl.wait_north = fl + 4                 # my flag[1]: raised by the north neighbour
                hl.signal_north = pb + st["flag_off"]  # its flag[0] ("from south")
            else:                # my row 0 -> south neighbour's row ny_south (index nyl_south + 1)
                if push:
                    for k in range(ncomp):
                        hl.south_ghost[k] = arr + k * comp + (st["nyl_south"] + 1) * nx * es
                hl.wait_south = fl
                hl.signal_south = pb + st["flag_off"] + 4
        return hl

    def run(self, st):
        import torch

        lib = self.lib
        h, flags, nyl, nb, nx, ncomp = st["h"], st["flags"], st["nyl"], st["nb"], st["nx"], st["ncomp"]
        n = int(self.spec.n_steps)
        es = st["es"]
        X, A, B, bar = st["X"], st["A"], st["B"], st["bar"]
        slab = {id(X): 0, id(A): 1, id(B): 2}
        base = self._epoch * (n + 1)  # flags only grow: value of (epoch, step s) is base + s + 1
        self._epoch += 1
        X[:, :, 1:nyl + 1].copy_(st["X0"])
        rows = st["rows"]

        def inner(t):
            return [(t[k].data_ptr() + nx * es, nx, rows * nx) for k in range(ncomp)]

        def plain(t):
            return [(t[k].data_ptr(), nx, nyl * nx) for k in range(ncomp)]

        stream = torch.cuda.current_stream(self.device).cuda_stream if self.device.type == "cuda" else 0
        if flags & _AREA_FLAG:
            lib.prepare(h, nb, inner(X), inner(B), stream)
            X, B = B, X
        # ghost rows of the prepared input; waits until the neighbours have finished the previous run
        lib.halo_push(h, nb, inner(X), self._halo(st, slab[id(X)], base, base + 1), stream)
        lib.cheb_step_halo(h, nb, 1, inner(X), None, inner(A), plain(bar),
                           self._halo(st, slab[id(A)], base + 1, base + 2), stream)
        T1, T2 = A, X
        for i in range(2, n + 1):
            D = T2
            hl = self._halo(st, slab[id(D)], base + i, base + i + 1, push=i < n)
            lib.cheb_step_halo(h, nb, i, inner(T1), inner(T2), inner(D), plain(bar), hl, stream)
            T2, T1 = T1, D
        return bar


class FusedBandedFilter(BandedFilter):
    """Latitude bands driven by the temporally blocked kernels: every rank keeps ``H`` ghost rows per side, runs ``H``
    Chebyshev steps per launch (``gcmf_cheb_fused`` on a band plan) and exchanges the ghost rows of ``T_{i+k-1}``
    and ``T_{i+k-2}`` **once per block** instead of once per step (north_star item 3: "per-step-block halo
    exchange").  Scalar FLUX / REGULAR5 operators (``H = 4``) on doubly periodic or tripolar grids (the top band
    folds onto itself inside the kernel, the bottom band has no southern neighbour; every band at least one tile
    = 32 rows high) and the vector operators (``H = 2``: the two-step row-streaming kernel; cfg5's 2160 x 4320
    C-grid field on 8 GPUs is eight bands of 270 rows); ``nx`` a multiple of the vector width.

    ``exchange="nccl"``: pack, grouped NCCL send/recv, unpack.  ``exchange="peer"``: the ghosted arrays live in
    symmetric memory; after a device-side barrier every rank *pulls* its ghost rows straight out of its
    neighbours' arrays over NVLink (four strided copies per block, no NCCL, no packing).  Two buffer pairs
    ping-pong between blocks, so one barrier per block also covers the write-after-read hazards.
    ``exchange="push"`` (vector operators): the exchange is **fused into the two-step kernel**
    (``gcmf_cheb_fused_halo``): the border row-bands store their two first / last rows of ``T_{i+1}`` and ``T_i``
    straight into the neighbours' ghost rows as they emit them and the last border CTA raises a flag in the
    neighbour's memory; the next block over there spins on that flag only in the CTAs that touch ghost rows.  One
    launch per two Chebyshev steps and rank, no NCCL call, no copy kernels, no host synchronisation; only the ghost
    rows of the input field are pulled once per call.  With one rank the band is its own neighbour (ordinary device
    memory): the same kernels and protocol on a single GPU.
    """

    def __init__(self, flt, rank, world, group=None, library=None, device=None, exchange="nccl"):
        super().__init__(flt, rank, world, group=group, library=library, device=device)
        if exchange not in ("nccl", "peer", "push"):
            raise ValueError("exchange must be 'nccl', 'peer' or 'push'")
        if exchange == "push" and self.lap.ncomp != 2:
            raise ValueError("exchange='push' is implemented for the vector operators (two-step kernel)")
        self.exchange = exchange if (self.world > 1 or exchange == "push") else "nccl"
        self._symm = None
        self._epoch = 0

    def close(self):
        import gc

        self._symm = None
        gc.collect()

    def _fused_plan(self, np_dtype, ny, nx):
        H = 4 if self.lap.ncomp == 1 else 2  # steps per block = ghost rows per side (scalar tiles: 4, vector rows: 2)
        h, keep, j0, j1, flags = self._plan(np_dtype, ny, nx, halo=H)
        if self.lib.fused_max_steps(h) < H:
            raise ValueError("this operator / band size has no fused kernel (scalar: band >= 32 rows, nx >= one tile; "
                             "vector: nx >= 232; nx a multiple of the 16-byte vector width)")
        return h, j0, j1, flags, H

    def stage(self, *fields):
        import torch

        lap = self.lap
        ncomp = lap.ncomp
        assert len(fields) == ncomp, f"this operator filters {ncomp} field(s) at a time"
        f0 = np.asarray(fields[0])
        ny, nx = f0.shape[-2:]
        np_dtype = lap.compute_dtype(f0.dtype if f0.dtype.kind == "f" else np.float64)
        h, j0, j1, flags, H = self._fused_plan(np_dtype, ny, nx)
        nyl = j1 - j0
        nb = int(np.prod(f0.shape[:-2])) if f0.ndim > 2 else 1
        tdt = torch.float32 if np_dtype == np.float32 else torch.float64
        bands = band_rows(ny, self.world)
        rows = max(b - a for a, b in bands) + 2 * H  # common allocation so that every rank has the same layout
        peers = None
        push = None
        n_flag = 64  # elements reserved behind the arrays for the flag / counter words of exchange="push"
        if self.exchange == "push" and self.world == 1:  # the band is its own neighbour: ordinary device memory
            n_arrays, slab = 6, ncomp * nb * rows * nx
            buf = torch.zeros(n_arrays * slab + n_flag, dtype=tdt, device=self.device)
            arrays = [buf[k * slab:(k + 1) * slab].view(ncomp, nb, rows, nx) for k in range(n_arrays)]
            push = dict(bases=[buf.data_ptr()], slab=slab, buf=buf, nyl_south=nyl)
            self._epoch = 0
        elif self.exchange in ("peer", "push"):
            import torch.distributed as dist
            import torch.distributed._symmetric_memory as symm_mem

            n_arrays, slab = 6, ncomp * nb * rows * nx
            if self._symm is None or self._symm[0] != (tdt, slab):
                buf = symm_mem.empty(n_arrays * slab + n_flag, dtype=tdt, device=self.device)
                hdl = symm_mem.rendezvous(buf, self.group if self.group is not None else dist.group.WORLD)
                self._symm = ((tdt, slab), buf, hdl)
            _, buf, hdl = self._symm
            # a neighbour's pull copies of the previous apply() may still be reading my arrays: meet first, then clear
            torch.cuda.synchronize(self.device)
            dist.barrier(group=self.group)
            buf.zero_()
            torch.cuda.synchronize(self.device)
            dist.barrier(group=self.group)
            arrays = [buf[k * slab:(k + 1) * slab].view(ncomp, nb, rows, nx) for k in range(n_arrays)]
            north, south = (self.rank + 1) % self.world, (self.rank - 1) % self.world
            peers = dict(hdl=hdl, nyl_south=bands[south][1] - bands[south][0],
                         north=[hdl.get_buffer(north, (ncomp, nb, rows, nx), tdt, k * slab) for k in range(n_arrays)],
                         south=[hdl.get_buffer(south, (ncomp, nb, rows, nx), tdt, k * slab) for k in range(n_arrays)])
            if self.exchange == "push":  # flags were cleared with the buffer, between two barriers: restart the epochs
                push = dict(bases=[int(p) for p in hdl.buffer_ptrs], slab=slab, buf=buf,
                            nyl_south=bands[south][1] - bands[south][0])
                self._epoch = 0
        else:
            arrays = [torch.zeros((ncomp, nb, rows, nx), dtype=tdt, device=self.device) for _ in range(6)]
        X0 = torch.stack([torch.as_tensor(np.ascontiguousarray(np.asarray(f).reshape((nb, ny, nx))[:, j0:j1]))
                          for f in fields]).to(device=self.device, dtype=tdt)
        return dict(h=h, flags=flags, j0=j0, j1=j1, nyl=nyl, nb=nb, nx=nx, ncomp=ncomp, H=H, rows=rows, X0=X0,
                    arrays=arrays, peers=peers, push=push,
                    bar=torch.empty((ncomp, nb, nyl, nx), dtype=tdt, device=self.device), batch_shape=f0.shape[:-2])

    def _push_halo(self, st, which, wait_value, signal_value, wait=True, push=True):
        """gcmf_halo describing where the border rows of array `which` go in the neighbours' copies of that array (and
        the flag words behind the arrays: [0] raised by the south neighbour, [1] by the north one, [2..3] counters)."""
        ps = st["push"]
        nx, rows, nb, ncomp, H = st["nx"], st["rows"], st["nb"], st["ncomp"], st["H"]
        es = st["arrays"][0].element_size()
        local = self.world == 1
        me = ps["bases"][0 if local else self.rank]
        north = ps["bases"][0 if local else (self.rank + 1) % self.world]
        south = ps["bases"][0 if local else (self.rank - 1) % self.world]
        flag_off = 6 * ps["slab"] * es
        comp = nb * rows * nx * es  # bytes of one component of a ghosted array
        hl = _cabi.Halo()
        hl.north_bstride = hl.south_bstride = rows * nx
        hl.counters = me + flag_off + 8
        hl.wait_value, hl.signal_value = wait_value & 0xFFFFFFFF, signal_value & 0xFFFFFFFF
        arr = which * ps["slab"] * es
        for k in range(ncomp):
            if push:  # my rows ny-2, ny-1 -> north's ghost rows -2, -1 (index 0, 1); my rows 0, 1 -> south's rows ny, ny+1
                hl.north_ghost[k] = north + arr + k * comp
                hl.south_ghost[k] = south + arr + k * comp + (H + ps["nyl_south"]) * nx * es
        if wait:
            hl.wait_north = me + flag_off + 4
            hl.wait_south = me + flag_off
        hl.signal_north = north + flag_off       # its flag[0]: "from south"
        hl.signal_south = south + flag_off + 4   # its flag[1]: "from north"
        return hl

    def _exchange_ghosts(self, st, idxs):
        """Fill the H ghost rows on both sides of the arrays `idxs` (owned rows sit at [H, H+nyl))."""
        H, nyl = st["H"], st["nyl"]
        fl = self.lap._planes.flags
        ring = bool(fl & _cabi.FLAG_WRAP_Y) and not (fl & _cabi.FLAG_CUT_S)  # tripolar grids do not wrap in y
        if st["peers"] is None:
            for k in idxs:  # NCCL send/recv on the rows actually in use
                self._exchange(st["arrays"][k][..., :nyl + 2 * H, :], ring, H)
            return
        pr = st["peers"]
        pr["hdl"].barrier(channel=0)  # every rank has finished writing its owned rows (and reading old ghosts)
        ns = pr["nyl_south"]
        for k in idxs:
            mine = st["arrays"][k]
            if ring or self.rank != 0:
                mine[..., 0:H, :].copy_(pr["south"][k][..., ns:ns + H, :])                # its top owned rows
            if ring or self.rank != self.world - 1:
                mine[..., H + nyl:2 * H + nyl, :].copy_(pr["north"][k][..., H:2 * H, :])  # its bottom owned rows

    def run(self, st):
        import torch

        lib = self.lib
        h, flags, nyl, nb, nx, H, rows = st["h"], st["flags"], st["nyl"], st["nb"], st["nx"], st["H"], st["rows"]
        n = int(self.spec.n_steps)
        A, bar = st["arrays"], st["bar"]
        es = A[0].element_size()

        ncomp = st["ncomp"]
        cache = st.setdefault("_ctypes", {})  # ctypes views built once per staging: the launches are host-bound on short bands

        def inner(k):  # the owned rows start H rows into the ghosted array
            if ("in", k) not in cache:
                cache[("in", k)] = lib.fields([(A[k][c].data_ptr() + H * nx * es, nx, rows * nx) for c in range(ncomp)])
            return cache[("in", k)]

        if "bar" not in cache:
            cache["bar"] = lib.fields([(bar[c].data_ptr(), nx, nyl * nx) for c in range(ncomp)])
        plain = cache["bar"]

        def halo(which, wait_value, signal_value, wait, push):
            key = ("halo", which, wait, push)
            if key not in cache:
                cache[key] = self._push_halo(st, which, 0, 0, wait=wait, push=push)
            hl = cache[key]
            hl.wait_value, hl.signal_value = wait_value & 0xFFFFFFFF, signal_value & 0xFFFFFFFF
            return hl
        stream = torch.cuda.current_stream(self.device).cuda_stream if self.device.type == "cuda" else 0
        x = 0
        A[0][:, :, H:H + nyl].copy_(st["X0"])
        if flags & _AREA_FLAG:  # x = f * area on the owned rows (kernels.py:100-101), then its ghosts
            lib.prepare(h, nb, inner(0), inner(1), stream)
            x = 1
        self._exchange_ghosts(st, [x])
        pairs = [(2, 3), (4, 5)]
        t1 = t2 = x
        cur, i = 0, 1
        if st["push"] is not None:
            # exchange fused into the kernels.  Flags only grow: block m of this call signals base + m and waits for
            # base + m - 1; the first block waits for nothing (the exchange of the input above is a barrier across
            # the ranks: everybody has finished the previous call).
            nblk = (n + 1) // 2
            base = self._epoch * (nblk + 1)
            self._epoch += 1
            m = 1
            while i <= n:
                kk = min(H, n - i + 1)
                o1, o2 = pairs[cur]
                if kk == 2:
                    last = i + 1 == n
                    h1 = halo(o1, base + m - 1, base + m, m > 1, not last)
                    h2 = halo(o2, base + m - 1, base + m, m > 1, not last)
                    lib.cheb_fused_halo(h, nb, i, 2, inner(t1), inner(t2), inner(o1), inner(o2), plain, h1, h2, stream)
                else:  # odd step count: the one-step LAST kernel, waiting for the ghost rows the previous block pushed
                    hl = halo(o1, base + m - 1, base + m, m > 1, False)
                    lib.cheb_step_halo(h, nb, i, inner(t1), inner(t2), inner(o1), plain, hl, stream)
                t1, t2 = o1, o2
                cur ^= 1
                i += kk
                m += 1
            return bar
        while i <= n:
            kk = min(H, n - i + 1)
            o1, o2 = pairs[cur]
            lib.cheb_fused(h, nb, i, kk, inner(t1), inner(t2), inner(o1), inner(o2), plain, stream)
            if i + kk <= n:  # the next block reads the ghost rows of both carried fields
                self._exchange_ghosts(st, [o1, o2])
            t1, t2 = o1, o2
            cur ^= 1
            i += kk
        return bar
