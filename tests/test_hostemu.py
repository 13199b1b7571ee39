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
