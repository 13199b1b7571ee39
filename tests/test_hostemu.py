"""CPU checks of the CUDA source through the host emulator: the stencil code of libgcmf.so
(gcmf_stencils.cuh), the C-ABI sequencing (gcmf.cu) and the host precombination
(gcm_filters_b200/kernels.py) against the oracle, for all 11 grid types."""
import numpy as np
import pytest

from gcm_filters_b200 import FilterShape, GridType
from gcm_filters_b200.filter import _compute_filter_spec, _compute_n_steps_default, _shift_scale
from gcm_filters_b200.kernels import ALL_KERNELS
from oracle import fixtures, np_oracle

from conftest import rel_l2
from hostemu_util import EmuPlan

ALL_GRIDS = fixtures.SCALAR_GRIDS + fixtures.VECTOR_GRIDS
BIT_EXACT = {"REGULAR", "REGULAR_AREA_WEIGHTED", "REGULAR_WITH_LAND", "REGULAR_WITH_LAND_AREA_WEIGHTED",
             "TRIPOLAR_REGULAR_WITH_LAND_AREA_WEIGHTED", "VECTOR_B_GRID"}


def vec_args(g, gv, fa):
    fa = dict(fa)
    if g in fixtures.VECTOR_GRIDS:
        kx, ky = ("dxT", "dyT") if g == "VECTOR_C_GRID" else ("DXU", "DYU")
        dxm = float(min(gv[kx].min(), gv[ky].min()))
        fa["dx_min"] = dxm
        fa["filter_scale"] = fa["filter_scale"] * dxm
    return fa


def spec_for(g, gv, **fa):
    fa = vec_args(g, gv, fa)
    shape = FilterShape[fa.get("filter_shape", "GAUSSIAN")]
    n = fa.get("n_steps", 0)
    if n < 3:
        n = _compute_n_steps_default(2, shape, fa["filter_scale"], fa["dx_min"], np.pi)
    return fa, _compute_filter_spec(fa["filter_scale"], fa["dx_min"], shape, np.pi, 2, n)


@pytest.mark.parametrize("shape", [(37, 54), (32, 48)])
@pytest.mark.parametrize("g", ALL_GRIDS)
def test_laplacian_f64(g, shape):
    fields, gv = fixtures.fixture(g, shape)
    lap = ALL_KERNELS[GridType[g]](**gv)
    ref = np_oracle.laplacian(g, gv, *fields)
    ref = ref if isinstance(ref, tuple) else (ref,)
    got = EmuPlan(lap, np.float64, *shape).laplacian(fields)
    for a, b in zip(got, ref):
        if g in BIT_EXACT:
            assert np.array_equal(a, b)
        else:
            assert rel_l2(a, b) < 1e-14


@pytest.mark.parametrize("shape", [(37, 54), (32, 48)])
@pytest.mark.parametrize("g", ALL_GRIDS)
def test_filter_f64(g, shape):
    fields, gv = fixtures.fixture(g, shape)
    lap = ALL_KERNELS[GridType[g]](**gv)
    fa, spec = spec_for(g, gv, filter_scale=8.0, dx_min=1.0)
    ref = np_oracle.apply_filter(g, gv, fields, **fa)
    ref = ref if isinstance(ref, tuple) else (ref,)
    got = EmuPlan(lap, np.float64, *shape).filter(fields, spec.p, _shift_scale(spec, lap))
    for a, b in zip(got, ref):
        if g in BIT_EXACT:
            assert np.array_equal(a, b)
        else:
            assert rel_l2(a, b) < 1e-12  # north_star tolerance, fp64


@pytest.mark.parametrize("g", ALL_GRIDS)
def test_filter_f32(g):
    shape = (40, 64)
    fields, gv = fixtures.fixture(g, shape)
    lap = ALL_KERNELS[GridType[g]](**gv)
    fa, spec = spec_for(g, gv, filter_scale=6.0, dx_min=1.0, filter_shape="TAPER")
    ref = np_oracle.apply_filter(g, gv, fields, **fa)
    ref = ref if isinstance(ref, tuple) else (ref,)
    f32 = tuple(f.astype(np.float32) for f in fields)
    got = EmuPlan(lap, np.float32, *shape).filter(f32, spec.p, _shift_scale(spec, lap))
    for a, b in zip(got, ref):
        assert a.dtype == np.float32
        assert rel_l2(a, b) < 1e-5  # north_star tolerance, fp32


@pytest.mark.parametrize("g", ["REGULAR_WITH_LAND", "IRREGULAR_WITH_LAND", "TRIPOLAR_POP_WITH_LAND",
                               "TRIPOLAR_REGULAR_WITH_LAND_AREA_WEIGHTED", "MOM5U", "MOM5T"])
def test_nan_on_land_batched(g):
    shape = (48, 64)
    (f,), gv = fixtures.fixture(g, shape)
    fb = np.stack([f, f[::-1].copy(), f * f])
    fb[:, gv["wet_mask"] == 0] = np.nan
    lap = ALL_KERNELS[GridType[g]](**gv)
    fa, spec = spec_for(g, gv, filter_scale=8.0, dx_min=1.0)
    ref = np_oracle.apply_filter(g, gv, (fb,), **fa)
    (got,) = EmuPlan(lap, np.float64, *shape).filter((fb,), spec.p, _shift_scale(spec, lap))
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    assert rel_l2(got, ref) < 1e-12


def test_batched_planes():
    # N5: a (z, y, x) wet mask against a (t, z, y, x) field
    shape = (24, 32)
    rng = np.random.default_rng(3)
    mask = (rng.random((3,) + shape) > 0.3).astype(np.float64)
    f = rng.random((2, 3) + shape)
    lap = ALL_KERNELS[GridType.REGULAR_WITH_LAND](wet_mask=mask)
    ref = np_oracle.laplacian("REGULAR_WITH_LAND", {"wet_mask": mask}, f)
    (got,) = EmuPlan(lap, np.float64, *shape).laplacian((f,))
    assert np.array_equal(got, ref)


def test_c_abi_argument_checks():
    """Error paths of the C ABI (same code in libgcmf.so): status code + message, nothing is launched."""
    from gcm_filters_b200 import _cabi
    from hostemu_util import emu_library
    lib = emu_library()
    (f,), gv = fixtures.fixture("IRREGULAR_WITH_LAND", (40, 136))
    lap = ALL_KERNELS[GridType.IRREGULAR_WITH_LAND](**gv)
    plan = EmuPlan(lap, np.float64, 40, 136)
    x = np.ascontiguousarray(f)[None].copy()
    y = np.zeros_like(x)
    spec = [(x.ctypes.data, 136, 40 * 136)]
    out = [(y.ctypes.data, 136, 40 * 136)]
    with pytest.raises(_cabi.GcmfError, match="gcmf_plan_set_filter has not been called"):
        lib.filter(plan.h, 1, spec, out, 0, 0)
    lib.plan_set_filter(plan.h, [0.5, 0.3, 0.1, 0.05], 0.25)
    with pytest.raises(_cabi.GcmfError, match="workspace too small"):
        lib.filter(plan.h, 1, spec, out, 0, 0)
    ws = np.zeros(lib.workspace_bytes(plan.h, 1) + 256, dtype=np.uint8)
    base = ws.ctypes.data + (-ws.ctypes.data) % 256
    with pytest.raises(_cabi.GcmfError, match="must not alias"):
        lib.filter(plan.h, 1, spec, spec, base, ws.size - 256)
    with pytest.raises(_cabi.GcmfError, match="outside 1"):
        lib.cheb_step(plan.h, 1, 9, spec, spec, out, out)
    with pytest.raises(_cabi.GcmfError, match="must lie inside"):
        lib.cheb_fused(plan.h, 1, 2, 4, spec, spec, out, out, out)
    with pytest.raises(_cabi.GcmfError, match="n_steps must be >= 2"):
        lib.plan_set_filter(plan.h, [1.0, 0.5], 0.25)
    with pytest.raises(_cabi.GcmfError, match="steps_per_block"):
        lib.set_steps_per_block(plan.h, 9)
    with pytest.raises(_cabi.GcmfError, match="grid .* too small"):
        lib.plan_create(_cabi.OP_FLUX, _cabi.GCMF_F64, 0, 8, _cabi.FLAG_WRAP_Y, 0)
    band = lib.plan_create(_cabi.OP_REGULAR5, _cabi.GCMF_F64, 40, 136, 0, 0)  # no WRAP_Y: a latitude band
    lib.plan_set_filter(band, [0.5, 0.3, 0.1, 0.05], 0.25)
    with pytest.raises(_cabi.GcmfError, match="latitude band"):
        lib.filter(band, 1, spec, out, base, ws.size - 256)
    lib.plan_destroy(band)


@pytest.mark.parametrize("shape", [(33, 40), (40, 136)])  # the second shape takes the fused path for REGULAR*
@pytest.mark.parametrize("g", ["REGULAR", "REGULAR_AREA_WEIGHTED", "VECTOR_C_GRID", "VECTOR_B_GRID"])
def test_nan_semantics_of_unmasked_and_vector_operators(g, shape):
    """SURVEY note N1: REGULAR has no nan_to_num (NaNs spread through the stencil), the vector operators zero
    NaNs inside the Laplacian but keep them in the `-x` term of the recurrence."""
    shape = (33, 40)
    fields, gv = fixtures.fixture(g, shape)
    fields = tuple(f.copy() for f in fields)
    fields[0][7, 9] = np.nan
    fields[-1][20, 31] = np.nan
    lap = ALL_KERNELS[GridType[g]](**gv)
    fa, spec = spec_for(g, gv, filter_scale=4.0, dx_min=1.0)
    ref = np_oracle.apply_filter(g, gv, fields, **fa)
    ref = ref if isinstance(ref, tuple) else (ref,)
    got = EmuPlan(lap, np.float64, *shape).filter(fields, spec.p, _shift_scale(spec, lap))
    for a, b in zip(got, ref):
        assert np.array_equal(np.isnan(a), np.isnan(b))
        assert np.isnan(b).sum() >= 1
        if g in BIT_EXACT:
            assert np.array_equal(a, b, equal_nan=True)
        else:
            assert rel_l2(a, b) < 1e-12


@pytest.mark.parametrize("shape", [(41, 48), (24, 37), (9, 20)])
@pytest.mark.parametrize("g", ["IRREGULAR_WITH_LAND", "VECTOR_C_GRID", "VECTOR_B_GRID",
                               "REGULAR_WITH_LAND_AREA_WEIGHTED", "TRIPOLAR_POP_WITH_LAND"])
def test_peer_banded_filter_single_rank(g, shape):
    """gcmf_halo_push / gcmf_cheb_step_halo (ghost rows stored by the step kernels, flag-synchronised) with one
    rank that is its own north and south neighbour, buffers in ordinary memory: the bookkeeping of
    PeerBandedFilter (slab layout, ghost-row addresses, growing flag values across two runs) against the oracle.
    The multi-GPU version of this test is tests/test_gpu_multi.py."""
    from gcm_filters_b200 import Filter, FilterShape
    from gcm_filters_b200.scheduler import PeerBandedFilter
    from hostemu_util import emu_library
    if g.startswith("TRIPOLAR") and shape[1] % 2:
        pytest.skip("the fold pairs column i with nx-1-i")
    fields, gv = fixtures.fixture(g, shape)
    fields = tuple(np.stack([f, f * f]) for f in fields)
    fa = dict(filter_scale=6.0, dx_min=1.0)
    if g in fixtures.VECTOR_GRIDS:
        kx, ky = ("dxT", "dyT") if g == "VECTOR_C_GRID" else ("DXU", "DYU")
        dxm = float(min(gv[kx].min(), gv[ky].min()))
        fa = dict(filter_scale=6.0 * dxm, dx_min=dxm)
    flt = Filter(grid_type=GridType[g], grid_vars=gv, filter_shape=FilterShape.GAUSSIAN, **fa)
    pf = PeerBandedFilter(flt, 0, 1, library=emu_library(), device="cpu")
    for _ in range(2):  # the second run re-uses the buffers with a new flag epoch
        outs, (j0, j1) = pf.apply(*fields)
    assert (j0, j1) == (0, shape[0])
    ref = np_oracle.apply_filter(g, gv, fields, **fa)
    ref = ref if isinstance(ref, tuple) else (ref,)
    for o, r in zip(outs, ref):
        assert rel_l2(o, r) < 1e-12


@pytest.mark.parametrize("shape,scale", [((41, 48), 6.0), ((24, 36), 7.0), ((9, 20), 6.0)])
@pytest.mark.parametrize("g", ["VECTOR_C_GRID", "VECTOR_B_GRID"])
def test_fused_banded_push_single_rank(g, shape, scale):
    """FusedBandedFilter(exchange="push") -- gcmf_cheb_fused_halo: two-step blocks whose border rows of T_{i+1} and T_i
    go straight into the neighbours' ghost rows -- with one rank that is its own north and south neighbour, buffers in
    ordinary memory: the host side of the protocol (addresses of the two ghost rows per side in the six rotating arrays,
    flag words behind them, growing flag values over several runs on one staging, the trailing one-step launch of an
    odd step count) against the oracle.  The emulator copies the border rows and raises the flags after each block;
    the device kernel and the multi-GPU protocol are tests/test_gpu_parity.py / tests/test_gpu_multi.py."""
    from gcm_filters_b200 import Filter, FilterShape
    from gcm_filters_b200.scheduler import FusedBandedFilter
    from hostemu_util import emu_library
    fields, gv = fixtures.fixture(g, shape)
    fields = tuple(np.stack([f, 1.0 - f * f]) for f in fields)
    kx, ky = ("dxT", "dyT") if g == "VECTOR_C_GRID" else ("DXU", "DYU")
    dxm = float(min(gv[kx].min(), gv[ky].min()))
    fa = dict(filter_scale=scale * dxm, dx_min=dxm)
    flt = Filter(grid_type=GridType[g], grid_vars=gv, filter_shape=FilterShape.GAUSSIAN, **fa)
    assert flt.n_steps == (7 if scale == 6.0 else 8)
    fbf = FusedBandedFilter(flt, 0, 1, library=emu_library(), device="cpu", exchange="push")
    assert fbf.exchange == "push"
    st = fbf.stage(*fields)
    for _ in range(3):
        bar = fbf.run(st)
    ref = np_oracle.apply_filter(g, gv, fields, **fa)
    for k, r in enumerate(ref):
        assert rel_l2(bar[k].numpy(), r) < 1e-12
    # the same through apply() (fresh staging, flags restart)
    outs, (j0, j1) = fbf.apply(*fields)
    assert (j0, j1) == (0, shape[0])
    for o, r in zip(outs, ref):
        assert rel_l2(o, r) < 1e-12
