"""Filter.apply / apply_to_vector on xarray objects (reference tests/test_filter.py:103-218, 256-290),
through real xarray where installed, else through tests/xr_shim.py."""
import numpy as np
import pytest

import xr_shim
from gcm_filters_b200 import Filter, FilterShape, GridType
from oracle import fixtures, np_oracle

from conftest import rel_l2


@pytest.fixture()
def xr():
    mod = xr_shim.install()
    yield mod
    xr_shim.uninstall()


def test_dataset_without_matching_dims_warns(xr):
    # reference tests/test_filter.py:207-218: no compute happens, so this runs without a GPU
    ds = xr.Dataset({"a": xr.DataArray(np.zeros((4, 5)), dims=["t", "z"])})
    flt = Filter(filter_scale=4.0, dx_min=1.0, grid_type=GridType.REGULAR)
    with pytest.warns(UserWarning, match=r".* nothing was filtered."):
        out = flt.apply(ds, dims=["y", "x"])
    assert np.array_equal(out["a"].data, ds["a"].data)


def test_grid_ds_is_a_dataset_for_dataarray_grid_vars(xr):
    (f,), gv = fixtures.fixture("REGULAR_WITH_LAND", (16, 24))
    gvx = {k: xr.DataArray(v, dims=["y", "x"]) for k, v in gv.items()}
    flt = Filter(filter_scale=4.0, dx_min=1.0, grid_type=GridType.REGULAR_WITH_LAND, grid_vars=gvx)
    assert isinstance(flt.grid_ds, xr.Dataset) and "grid_vars" not in repr(flt)


@pytest.mark.gpu
def test_dataarray_with_batch_dims_in_any_order(xr):
    """Core dims are named, not positional: a (y, time, x, depth) array filters over (y, x)."""
    (f,), gv = fixtures.fixture("IRREGULAR_WITH_LAND", (48, 64))
    rng = np.random.default_rng(0)
    data = f[None, None] * (1 + 0.1 * rng.standard_normal((3, 2, 1, 1)))  # (time, depth, y, x)
    gvx = {k: xr.DataArray(v, dims=["y", "x"]) for k, v in gv.items()}
    flt = Filter(filter_scale=6.0, dx_min=1.0, grid_type=GridType.IRREGULAR_WITH_LAND, grid_vars=gvx)
    ref = np_oracle.apply_filter("IRREGULAR_WITH_LAND", gv, (data,), filter_scale=6.0, dx_min=1.0)
    da = xr.DataArray(np.transpose(data, (2, 0, 3, 1)), dims=["y", "time", "x", "depth"])
    out = flt.apply(da, dims=["y", "x"])
    assert tuple(out.dims) == ("time", "depth", "y", "x")
    assert rel_l2(np.asarray(out.data), ref) < 1e-12


@pytest.mark.gpu
def test_dataset_semantics(xr):
    # reference tests/test_filter.py:172-205: only variables holding both dims are filtered; means preserved
    (f,), gv = fixtures.fixture("REGULAR", (32, 48))
    ds = xr.Dataset({
        "spatial": xr.DataArray(f, dims=["y", "x"]),
        "spacetime": xr.DataArray(np.stack([f, 2 * f]), dims=["time", "y", "x"]),
        "temporal": xr.DataArray(np.arange(5.0), dims=["time5"]),
    })
    flt = Filter(filter_scale=4.0, dx_min=1.0, grid_type=GridType.REGULAR)
    out = flt.apply(ds, dims=["y", "x"])
    assert np.array_equal(out["temporal"].data, ds["temporal"].data)
    ref = np_oracle.apply_filter("REGULAR", {}, (f,), filter_scale=4.0, dx_min=1.0)
    assert np.array_equal(np.asarray(out["spatial"].data), ref)
    assert np.array_equal(np.asarray(out["spacetime"].data)[1], np_oracle.apply_filter("REGULAR", {}, (2 * f,), filter_scale=4.0, dx_min=1.0))
    np.testing.assert_allclose(np.asarray(out["spatial"].data).mean(), f.mean(), rtol=1e-12)
    assert np.array_equal(ds["spatial"].data, f)  # the input dataset is untouched (deep copy)


@pytest.mark.gpu
def test_vector_filter_on_dataarrays(xr):
    # reference tests/test_filter.py:256-290 (Taper, n_steps 10)
    (u, v), gv = fixtures.fixture("VECTOR_C_GRID", (48, 64))
    dxm = float(min(gv["dxT"].min(), gv["dyT"].min()))
    gvx = {k: xr.DataArray(a, dims=["y", "x"]) for k, a in gv.items()}
    flt = Filter(filter_scale=5.0 * dxm, dx_min=dxm, filter_shape=FilterShape.TAPER, n_steps=10,
                 grid_type=GridType.VECTOR_C_GRID, grid_vars=gvx)
    fu, fv = flt.apply_to_vector(xr.DataArray(u, dims=["y", "x"]), xr.DataArray(v, dims=["y", "x"]), dims=["y", "x"])
    ru, rv = np_oracle.apply_filter("VECTOR_C_GRID", gv, (u, v), filter_scale=5.0 * dxm, dx_min=dxm,
                                    filter_shape="TAPER", n_steps=10)
    assert rel_l2(np.asarray(fu.data), ru) < 1e-12 and rel_l2(np.asarray(fv.data), rv) < 1e-12
    with pytest.raises(ValueError, match=r".* is a vector Laplacian.*"):
        flt.apply(xr.DataArray(u, dims=["y", "x"]), dims=["y", "x"])
