"""Multi-GPU tests (need >= 2 CUDA devices; skipped otherwise): NCCL halo exchange of the latitude-band
decomposition and batch sharding, against the single-GPU result."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, g, shape, q):
    import torch
    import torch.distributed as dist
    from gcm_filters_b200 import Filter, FilterShape, GridType
    from gcm_filters_b200.scheduler import BandedFilter, FusedBandedFilter, PeerBandedFilter, apply_batch_sharded
    from oracle import fixtures

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        fields, gv = fixtures.fixture(g, shape)
        fields = tuple(np.stack([f, f * f, 1 - f]) for f in fields)
        fa = dict(filter_scale=6.0, dx_min=1.0)
        if g in fixtures.VECTOR_GRIDS:
            dxm = float(min(gv["dxT"].min(), gv["dyT"].min()))
            fa = dict(filter_scale=6.0 * dxm, dx_min=dxm)
        flt = Filter(grid_type=GridType[g], grid_vars=gv, filter_shape=FilterShape.GAUSSIAN, **fa)
        # reference = the one-step kernels on one GPU (the fused path differs in the last bits next to a tripolar
        # fold, where mirrored cells sum their fluxes in the opposite order)
        from gcm_filters_b200 import engine
        engine.set_steps_per_block(1)
        if len(fields) == 2:
            single = flt.apply_to_vector(fields[0], fields[1], dims=["y", "x"])
        else:
            single = (flt.apply(fields[0], dims=["y", "x"]),)
        outs, (j0, j1) = BandedFilter(flt, rank, world).apply(*fields)
        ok = all(np.array_equal(o, s[..., j0:j1, :], equal_nan=True) for o, s in zip(outs, single))
        # the same decomposition with the ghost-row exchange fused into the kernels (peer memory over NVLink)
        pbf = PeerBandedFilter(flt, rank, world)
        for _ in range(3):  # several epochs: the flags only grow
            pouts, (pj0, pj1) = pbf.apply(*fields)
        ok = ok and (pj0, pj1) == (j0, j1) and all(np.array_equal(o, s[..., j0:j1, :], equal_nan=True)
                                                   for o, s in zip(pouts, single))
        pbf.close()
        # temporal blocking on bands: 4 (scalar) / 2 (vector) ghost rows, one exchange per block of that many steps
        for exch in (("nccl", "peer", "push") if len(fields) == 2 else ("nccl", "peer")):
            fbf = FusedBandedFilter(flt, rank, world, exchange=exch)
            for _ in range(2):
                fouts, (fj0, fj1) = fbf.apply(*fields)
            ok = ok and (fj0, fj1) == (j0, j1)
            for fo, sg in zip(fouts, single):
                ref_band = sg[..., j0:j1, :]
                same_nan = np.array_equal(np.isnan(fo), np.isnan(ref_band))
                wet = ~np.isnan(ref_band)
                err = np.linalg.norm(fo[wet] - ref_band[wet]) / np.linalg.norm(ref_band[wet])
                # bit-identical on periodic grids; next to a tripolar fold mirrored cells sum in the opposite order
                ok = ok and same_nan and (err < 1e-14 if g == "TRIPOLAR_POP_WITH_LAND" else err == 0.0)
            fbf.close()
        # batch sharding with an all-gather of the slabs
        full = (apply_batch_sharded(lambda a: flt.apply(a, None), fields[0], rank, world, gather=True)
                if len(fields) == 1 else None)
        ok2 = True if full is None else bool(np.array_equal(full, single[0], equal_nan=True))
        q.put((rank, bool(ok), ok2))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("g", ["IRREGULAR_WITH_LAND", "TRIPOLAR_POP_WITH_LAND", "VECTOR_C_GRID"])
def test_banded_and_sharded_match_single_gpu(g, world):
    """Bands of 48 (45) rows per GPU: BandedFilter (NCCL per step), PeerBandedFilter (flag-synchronised stores into
    peer memory, three epochs on reused symmetric buffers), FusedBandedFilter (4-step scalar / 2-step vector blocks)
    with both exchanges, batch sharding -- all against the single-GPU one-step kernels, bit for bit."""
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    shape = (45 * world, 240) if g == "VECTOR_C_GRID" else (48 * world, 264)  # 240 columns: one strip of the two-step kernel
    procs = [ctx.Process(target=_worker, args=(r, world, port, g, shape, q))
             for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    res = [q.get(timeout=10) for _ in range(world)]
    assert all(r[1] and r[2] for r in res), res


def test_in_process_sharding_over_all_devices():
    """set_devices("all"): Filter.apply on a host batch is cut into slabs, one per GPU of this process (host threads,
    no torchrun); same bits as the single-device call, for numpy in / numpy out and for pinned tensors with out=."""
    n = _ngpu()
    if n < 2:
        pytest.skip("needs 2 GPUs")
    import torch
    import gcm_filters_b200 as gf
    from oracle import fixtures
    (f,), gv = fixtures.fixture("IRREGULAR_WITH_LAND", (96, 264))
    rng = np.random.default_rng(4)
    fb = f[None] * (1 + 0.2 * rng.standard_normal((4 * n + 3, 1, 1)))
    fb[:, gv["wet_mask"] == 0] = np.nan
    flt = gf.Filter(filter_scale=8.0, dx_min=1.0, grid_type=gf.GridType.IRREGULAR_WITH_LAND, grid_vars=gv)
    single = flt.apply(fb, None)
    try:
        gf.set_devices("all")
        multi = flt.apply(fb, None)
        pin_in = torch.from_numpy(fb).pin_memory()
        pin_out = torch.empty_like(pin_in).pin_memory()
        flt.apply(pin_in, None, out=pin_out)
    finally:
        gf.set_devices(None)
    assert isinstance(multi, np.ndarray) and np.array_equal(multi, single, equal_nan=True)
    assert np.array_equal(pin_out.numpy(), single, equal_nan=True)
