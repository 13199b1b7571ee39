"""Multi-process (gloo, CPU) tests of the multi-GPU scheduling logic: batch slabs and the latitude-band
decomposition with ghost-row exchange.  Compute goes through the TEST-ONLY host emulator of the C ABI
(same sources as libgcmf.so); the comparison target is the CPU oracle on the undecomposed domain."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gcm_filters_b200 import Filter, FilterShape, GridType
from gcm_filters_b200.scheduler import BandedFilter, apply_batch_sharded, band_rows, batch_slabs
from oracle import fixtures, np_oracle

from hostemu_util import emu_library


def test_slab_arithmetic():
    assert batch_slabs(62, 8) == [(0, 8), (8, 16), (16, 24), (24, 32), (32, 40), (40, 48), (48, 55), (55, 62)]
    assert [b - a for a, b in batch_slabs(365, 8)] == [46, 46, 46, 46, 46, 45, 45, 45]
    assert batch_slabs(3, 4) == [(0, 1), (1, 2), (2, 3), (3, 3)]
    assert band_rows(2160, 8)[0] == (0, 270) and band_rows(2160, 8)[-1] == (1890, 2160)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _vec_args(g, gv, fa):
    fa = dict(fa)
    if g in fixtures.VECTOR_GRIDS:
        kx, ky = ("dxT", "dyT") if g == "VECTOR_C_GRID" else ("DXU", "DYU")
        dxm = float(min(gv[kx].min(), gv[ky].min()))
        fa["dx_min"] = dxm
        fa["filter_scale"] = fa["filter_scale"] * dxm
    return fa


def _band_worker(rank, world, port, g, shape, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        fields, gv = fixtures.fixture(g, shape)
        fields = tuple(np.stack([f, f * f]) for f in fields)  # two batch slices
        fa = _vec_args(g, gv, dict(filter_scale=6.0, dx_min=1.0))
        flt = Filter(grid_type=GridType[g], grid_vars=gv, filter_shape=FilterShape.GAUSSIAN, **fa)
        bf = BandedFilter(flt, rank, world, library=emu_library(), device="cpu")
        outs, (j0, j1) = bf.apply(*fields)
        ref = np_oracle.apply_filter(g, gv, fields, **fa)
        ref = ref if isinstance(ref, tuple) else (ref,)
        err = 0.0
        for o, r in zip(outs, ref):
            rb = r[..., j0:j1, :]
            assert np.array_equal(np.isnan(o), np.isnan(rb))
            ok = ~np.isnan(rb)
            err = max(err, float(np.linalg.norm(o[ok] - rb[ok]) / np.linalg.norm(rb[ok])))
        q.put((rank, j0, j1, err))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("g", ["IRREGULAR_WITH_LAND", "TRIPOLAR_POP_WITH_LAND", "VECTOR_C_GRID",
                               "REGULAR_WITH_LAND_AREA_WEIGHTED", "TRIPOLAR_REGULAR_WITH_LAND_AREA_WEIGHTED"])
def test_band_decomposition_matches_single_domain(g, world):
    emu_library()  # build once in the parent
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    shape = (41, 48)
    procs = [ctx.Process(target=_band_worker, args=(r, world, port, g, shape, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = sorted(q.get(timeout=10) for _ in range(world))
    assert [(r[1], r[2]) for r in res] == band_rows(shape[0], world)
    assert max(r[3] for r in res) < 1e-12


def _slab_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(5)
        field = rng.random((7, 3, 12, 16))
        full = apply_batch_sharded(lambda a: a * 2.0 + 1.0, field, rank, world, gather=True)
        q.put((rank, bool(np.array_equal(full, field * 2.0 + 1.0))))
    finally:
        dist.destroy_process_group()


def test_batch_sharding_gather():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    world = 2
    procs = [ctx.Process(target=_slab_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert all(ok for _, ok in (q.get(timeout=10) for _ in range(world)))


def _fused_band_worker(rank, world, port, g, shape, dtype, q):
    from gcm_filters_b200.scheduler import FusedBandedFilter
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        (f,), gv = fixtures.fixture(g, shape)
        fb = np.stack([f, f * f, 1 - f]).astype(dtype)
        if "wet_mask" in gv:
            fb[:, gv["wet_mask"] == 0] = np.nan
        gvt = {k: v.astype(dtype) for k, v in gv.items()}
        fa = dict(filter_scale=10.0, dx_min=1.0)
        flt = Filter(grid_type=GridType[g], grid_vars=gvt, filter_shape=FilterShape.GAUSSIAN, **fa)
        bf = FusedBandedFilter(flt, rank, world, library=emu_library(), device="cpu")
        for _ in range(2):
            outs, (j0, j1) = bf.apply(fb)
        ref = np_oracle.apply_filter(g, gv, (fb.astype(np.float64),), **fa)[..., j0:j1, :]
        assert np.array_equal(np.isnan(outs[0]), np.isnan(ref))
        ok = ~np.isnan(ref)
        err = float(np.linalg.norm(outs[0][ok] - ref[ok]) / np.linalg.norm(ref[ok]))
        q.put((rank, j0, j1, err))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("g,dtype,tol", [("IRREGULAR_WITH_LAND", np.float64, 1e-12), ("REGULAR_WITH_LAND", np.float64, 1e-15),
                                         ("REGULAR_WITH_LAND_AREA_WEIGHTED", np.float32, 1e-5),
                                         ("TRIPOLAR_POP_WITH_LAND", np.float64, 1e-12),
                                         ("TRIPOLAR_REGULAR_WITH_LAND_AREA_WEIGHTED", np.float64, 1e-12)])
@pytest.mark.parametrize("world", [2, 3])
def test_fused_band_decomposition(g, dtype, tol, world):
    """Temporal blocking on latitude bands: 4 ghost rows, one exchange per 4-step block."""
    emu_library()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    shape = (112, 264)
    procs = [ctx.Process(target=_fused_band_worker, args=(r, world, port, g, shape, dtype, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    res = sorted(q.get(timeout=10) for _ in range(world))
    assert [(r[1], r[2]) for r in res] == band_rows(shape[0], world)
    assert max(r[3] for r in res) < tol


def _fused_vector_band_worker(rank, world, port, g, shape, scale, q):
    from gcm_filters_b200.scheduler import FusedBandedFilter
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        (u, v), gv = fixtures.fixture(g, shape)
        fields = (np.stack([u, u * v]), np.stack([v, 1 - u]))  # two batch slices per component
        fa = _vec_args(g, gv, dict(filter_scale=scale, dx_min=1.0))
        flt = Filter(grid_type=GridType[g], grid_vars=gv, filter_shape=FilterShape.GAUSSIAN, **fa)
        bf = FusedBandedFilter(flt, rank, world, library=emu_library(), device="cpu")
        for _ in range(2):
            outs, (j0, j1) = bf.apply(*fields)
        ref = np_oracle.apply_filter(g, gv, fields, **fa)
        err = 0.0
        for o, r in zip(outs, ref):
            rb = r[..., j0:j1, :]
            err = max(err, float(np.linalg.norm(o - rb) / np.linalg.norm(rb)))
        q.put((rank, j0, j1, err, int(flt.n_steps)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("g", ["VECTOR_C_GRID", "VECTOR_B_GRID"])
@pytest.mark.parametrize("world,scale", [(2, 6.0), (3, 7.0)])
def test_fused_band_decomposition_vector(g, world, scale):
    """Two-step blocks of the vector operators on latitude bands: 2 ghost rows per side of the fields and of every
    coefficient plane, one exchange of the ghost rows of T_{i+1} and T_i per block, a trailing one-step launch when the
    step count is odd (scale 6: 7 steps, scale 7: 8).  The emulator runs a block on a band as step i on the band extended
    by one row per side followed by step i+1 on the band -- the data flow of the device kernel."""
    emu_library()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    shape = (45, 40)
    procs = [ctx.Process(target=_fused_vector_band_worker, args=(r, world, port, g, shape, scale, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    res = sorted(q.get(timeout=10) for _ in range(world))
    assert [(r[1], r[2]) for r in res] == band_rows(shape[0], world)
    assert res[0][4] == (7 if scale == 6.0 else 8)
    assert max(r[3] for r in res) < 1e-12
