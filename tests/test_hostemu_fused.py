"""CPU checks of the temporally blocked kernel (gcmf_fused.cuh) through the host emulator: the
fused path must be bit-identical to the one-step path, and within tolerance of the oracle."""
import numpy as np
import pytest

from gcm_filters_b200 import FilterShape, GridType
from gcm_filters_b200.filter import _compute_filter_spec, _shift_scale
from gcm_filters_b200.kernels import ALL_KERNELS
from oracle import fixtures, np_oracle

from conftest import rel_l2
from hostemu_util import EmuPlan, emu_set_steps_per_block


def _case(g, shape, nb, dtype, n_steps, seed=0):
    (f,), gv = fixtures.fixture(g, shape)
    rng = np.random.default_rng(seed)
    fb = np.stack([f * (1 + 0.1 * k) + 0.05 * rng.standard_normal(shape) for k in range(nb)])
    fb[:, gv["wet_mask"] == 0] = np.nan
    lap = ALL_KERNELS[GridType[g]](**gv)
    spec = _compute_filter_spec(8.0, 1.0, FilterShape.GAUSSIAN, np.pi, 2, n_steps)
    return fb.astype(dtype), gv, lap, spec


@pytest.mark.parametrize("g", ["IRREGULAR_WITH_LAND", "MOM5T"])
@pytest.mark.parametrize("shape,nb,n_steps", [((70, 250), 3, 11), ((36, 128), 1, 7), ((100, 136), 2, 4)])
def test_fused_equals_unfused_f64(g, shape, nb, n_steps):
    fb, gv, lap, spec = _case(g, shape, nb, np.float64, n_steps)
    c = _shift_scale(spec, lap)
    fused = EmuPlan(lap, np.float64, *shape)
    assert fused.lib.fused_max_steps(fused.h) == 4
    n0 = fused.lib.launch_count()
    (a,) = fused.filter((fb,), spec.p, c)
    launches = fused.lib.launch_count() - n0
    assert launches == -(-n_steps // 4)  # the whole recurrence in ceil(n/4) fused launches
    plain = EmuPlan(lap, np.float64, *shape)
    emu_set_steps_per_block(plain, 1)
    (b,) = plain.filter((fb,), spec.p, c)
    assert np.array_equal(a, b, equal_nan=True)
    ref = np_oracle.run_recurrence(np_oracle.make_operator(g, gv), np_oracle.FilterSpec(*spec), (fb,))
    assert rel_l2(a, ref) < 1e-12


@pytest.mark.parametrize("k", [2, 3])
def test_fused_block_cap(k):
    fb, gv, lap, spec = _case("IRREGULAR_WITH_LAND", (64, 200), 2, np.float64, 9)
    c = _shift_scale(spec, lap)
    capped = EmuPlan(lap, np.float64, 64, 200)
    emu_set_steps_per_block(capped, k)
    (a,) = capped.filter((fb,), spec.p, c)
    plain = EmuPlan(lap, np.float64, 64, 200)
    emu_set_steps_per_block(plain, 1)
    (b,) = plain.filter((fb,), spec.p, c)
    assert np.array_equal(a, b, equal_nan=True)


def test_fused_f32():
    shape = (60, 300)
    fb, gv, lap, spec = _case("IRREGULAR_WITH_LAND", shape, 2, np.float32, 10)
    c = _shift_scale(spec, lap)
    fused = EmuPlan(lap, np.float32, *shape)
    assert fused.lib.fused_max_steps(fused.h) == 4
    (a,) = fused.filter((fb,), spec.p, c)
    plain = EmuPlan(lap, np.float32, *shape)
    emu_set_steps_per_block(plain, 1)
    (b,) = plain.filter((fb,), spec.p, c)
    assert np.array_equal(a, b, equal_nan=True)
    ref = np_oracle.run_recurrence(np_oracle.make_operator("IRREGULAR_WITH_LAND", gv), np_oracle.FilterSpec(*spec),
                                   (fb.astype(np.float64),))
    assert rel_l2(a, ref) < 1e-5


def test_fused_not_offered_for_small_grids():
    (f,), gv = fixtures.fixture("IRREGULAR_WITH_LAND", (30, 100))
    lap = ALL_KERNELS[GridType.IRREGULAR_WITH_LAND](**gv)
    pl = EmuPlan(lap, np.float64, 30, 100)
    assert pl.lib.fused_max_steps(pl.h) == 0


@pytest.mark.parametrize("g", ["TRIPOLAR_POP_WITH_LAND", "TRIPOLAR_REGULAR_WITH_LAND_AREA_WEIGHTED"])
@pytest.mark.parametrize("dtype,shape,n_steps", [(np.float64, (70, 250), 11), (np.float64, (33, 128), 6),
                                                  (np.float32, (50, 300), 9)])
def test_fused_tripolar_fold(g, dtype, shape, n_steps):
    """Tiles that reach across the tripolar fold evolve virtual (mirrored) halo rows."""
    (f,), gv = fixtures.fixture(g, shape)
    rng = np.random.default_rng(4)
    fb = np.stack([f + 0.1 * rng.standard_normal(shape), f * f])
    fb[:, gv["wet_mask"] == 0] = np.nan
    lap = ALL_KERNELS[GridType[g]](**gv)
    spec = _compute_filter_spec(8.0, 1.0, FilterShape.GAUSSIAN, np.pi, 2, n_steps)
    c = _shift_scale(spec, lap)
    fused = EmuPlan(lap, dtype, *shape)
    assert fused.lib.fused_max_steps(fused.h) == 4
    (a,) = fused.filter((fb.astype(dtype),), spec.p, c)
    plain = EmuPlan(lap, dtype, *shape)
    emu_set_steps_per_block(plain, 1)
    (b,) = plain.filter((fb.astype(dtype),), spec.p, c)
    assert np.array_equal(np.isnan(a), np.isnan(b))
    # mirrored cells sum their E/W and N/S fluxes in the opposite order: equal to rounding, not bit for bit
    assert rel_l2(a, b) < (1e-14 if dtype == np.float64 else 1e-6)
    ref = np_oracle.run_recurrence(np_oracle.make_operator(g, gv), np_oracle.FilterSpec(*spec), (fb,))
    assert rel_l2(a, ref) < (1e-12 if dtype == np.float64 else 1e-5)
    top = rel_l2(a[..., -6:, :], ref[..., -6:, :])  # the rows next to the fold
    assert top < (1e-12 if dtype == np.float64 else 1e-5)


@pytest.mark.parametrize("g", ["REGULAR", "REGULAR_WITH_LAND", "REGULAR_WITH_LAND_AREA_WEIGHTED", "REGULAR_AREA_WEIGHTED"])
@pytest.mark.parametrize("dtype,shape", [(np.float64, (70, 250)), (np.float32, (40, 300))])
def test_fused_regular5(g, dtype, shape):
    (f,), gv = fixtures.fixture(g, shape)
    fb = np.stack([f, f * f, 1 - f])
    if "wet_mask" in gv:
        fb[:, gv["wet_mask"] == 0] = np.nan
    lap = ALL_KERNELS[GridType[g]](**gv)
    spec = _compute_filter_spec(8.0, 1.0, FilterShape.GAUSSIAN, np.pi, 2, 11)
    c = _shift_scale(spec, lap)
    fused = EmuPlan(lap, dtype, *shape)
    assert fused.lib.fused_max_steps(fused.h) == 4
    n0 = fused.lib.launch_count()
    (a,) = fused.filter((fb.astype(dtype),), spec.p, c)
    assert fused.lib.launch_count() - n0 == 3 + (1 if "AREA" in g else 0)  # ceil(11/4) (+ prepare)
    plain = EmuPlan(lap, dtype, *shape)
    emu_set_steps_per_block(plain, 1)
    (b,) = plain.filter((fb.astype(dtype),), spec.p, c)
    assert np.array_equal(a, b, equal_nan=True)
    ref = np_oracle.run_recurrence(np_oracle.make_operator(g, gv), np_oracle.FilterSpec(*spec), (fb,))
    if dtype == np.float64:
        assert np.array_equal(a, ref, equal_nan=True)  # REGULAR5 family: bit-identical to the reference arithmetic
    else:
        assert rel_l2(a, ref) < 1e-5


@pytest.mark.parametrize("g,dtype", [("IRREGULAR_WITH_LAND", np.float64), ("REGULAR_WITH_LAND", np.float32),
                                     ("TRIPOLAR_POP_WITH_LAND", np.float64)])
@pytest.mark.parametrize("levels", [2, 3, 7])
def test_fused_level_slabs(g, dtype, levels, monkeypatch):
    """One CTA loops over a slab of levels (landing tiles refilled while the steps run, mbarrier parity flips,
    progress counters keep growing).  Small grids never get slabs longer than one level on their own, so the
    slab length is forced."""
    monkeypatch.setenv("GCMF_FUSED_LEVELS_PER_CTA", str(levels))
    shape = (40, 264)
    (f,), gv = fixtures.fixture(g, shape)
    rng = np.random.default_rng(levels)
    fb = f[None] * (1 + 0.3 * rng.standard_normal((7, 1, 1)))
    fb[:, gv["wet_mask"] == 0] = np.nan
    lap = ALL_KERNELS[GridType[g]](**gv)
    spec = _compute_filter_spec(8.0, 1.0, FilterShape.GAUSSIAN, np.pi, 2, 9)
    c = _shift_scale(spec, lap)
    fused = EmuPlan(lap, dtype, *shape)
    (a,) = fused.filter((fb.astype(dtype),), spec.p, c)
    monkeypatch.delenv("GCMF_FUSED_LEVELS_PER_CTA")
    plain = EmuPlan(lap, dtype, *shape)
    emu_set_steps_per_block(plain, 1)
    (b,) = plain.filter((fb.astype(dtype),), spec.p, c)
    if g.startswith("TRIPOLAR"):
        assert np.array_equal(np.isnan(a), np.isnan(b)) and rel_l2(a, b) < 1e-14
    else:
        assert np.array_equal(a, b, equal_nan=True)


@pytest.mark.parametrize("g", ["VECTOR_C_GRID", "VECTOR_B_GRID"])
@pytest.mark.parametrize("shape,nb,n_steps", [((40, 64), 2, 6), ((21, 30), 1, 7), ((36, 50), 3, 3)])
def test_vector_two_step_blocks_equal_one_step(g, shape, nb, n_steps):
    """The blocked control flow of the vector recurrence (gcmf_filter -> gcmf_cheb_fused with k = 2 on four ping-ponging
    workspace fields, a trailing one-step launch for odd step counts) through the emulator, which runs a block as its
    two one-step launches: identical bits, half the launches (the device kernel is pinned against the same one-step
    kernels by the GPU suite)."""
    (u, v), gv = fixtures.fixture(g, shape)
    us = np.stack([u * (1 + 0.1 * k) for k in range(nb)])
    vs = np.stack([v - 0.2 * k * u for k in range(nb)])
    us[0, 3:5, 7:9] = np.nan
    lap = ALL_KERNELS[GridType[g]](**gv)
    kx, ky = ("dxT", "dyT") if g == "VECTOR_C_GRID" else ("DXU", "DYU")
    dxm = float(min(gv[kx].min(), gv[ky].min()))
    spec = _compute_filter_spec(6.0 * dxm, dxm, FilterShape.GAUSSIAN, np.pi, 2, n_steps)
    c = _shift_scale(spec, lap)
    blocked = EmuPlan(lap, np.float64, *shape)
    assert blocked.lib.fused_max_steps(blocked.h) == 2
    n0 = blocked.lib.launch_count()
    a = blocked.filter((us, vs), spec.p, c)
    assert blocked.lib.launch_count() - n0 == (n_steps + 1) // 2
    plain = EmuPlan(lap, np.float64, *shape)
    emu_set_steps_per_block(plain, 1)
    n0 = plain.lib.launch_count()
    b = plain.filter((us, vs), spec.p, c)
    assert plain.lib.launch_count() - n0 == n_steps
    for x, y in zip(a, b):
        assert np.array_equal(x, y, equal_nan=True)
    ref = np_oracle.run_recurrence(np_oracle.make_operator(g, gv), np_oracle.FilterSpec(*spec), (us, vs))
    for x, r in zip(a, ref):
        assert rel_l2(x, r) < 1e-12
