"""The reference's own test-suite (tests/test_filter.py, tests/test_kernels.py), test by test, against this
implementation's public API -- same fixtures (oracle/fixtures.py restates tests/conftest.py seed for seed), same
assertions, same tolerances.

Every test runs on two backends:
  * ``cuda``    (marked ``gpu``): the product -- Filter / Laplacian -> engine -> C ABI -> sm_100a kernels;
  * ``hostemu`` (CPU suite): the same Python host code and the same CUDA source compiled for the host
    (tests/emu_backend.py).  It checks everything but the device execution itself, and it proves that the
    assertions below are satisfiable before the GPU run.
The xarray branch of Filter.apply is exercised through real xarray where installed, else tests/xr_shim.py.
"""
import copy

import numpy as np
import pytest

import emu_backend
import xr_shim
from gcm_filters_b200 import Filter, FilterShape, GridType, required_grid_vars
from gcm_filters_b200.kernels import ALL_KERNELS, AreaWeightedMixin
from oracle import fixtures

BACKENDS = [pytest.param("hostemu"), pytest.param("cuda", marks=pytest.mark.gpu)]

# reference conftest.py:62-76
SCALAR_GRIDS = ["REGULAR", "REGULAR_AREA_WEIGHTED", "REGULAR_WITH_LAND", "REGULAR_WITH_LAND_AREA_WEIGHTED",
                "IRREGULAR_WITH_LAND", "TRIPOLAR_REGULAR_WITH_LAND_AREA_WEIGHTED", "TRIPOLAR_POP_WITH_LAND"]
IRREGULAR_GRIDS = ["IRREGULAR_WITH_LAND", "TRIPOLAR_POP_WITH_LAND"]
TRIPOLAR_GRIDS = ["TRIPOLAR_REGULAR_WITH_LAND_AREA_WEIGHTED", "TRIPOLAR_POP_WITH_LAND"]
VECTOR_GRIDS = ["VECTOR_C_GRID", "VECTOR_B_GRID"]
AREA_WEIGHTED_REGULAR = ["REGULAR_AREA_WEIGHTED", "REGULAR_WITH_LAND_AREA_WEIGHTED",
                         "TRIPOLAR_REGULAR_WITH_LAND_AREA_WEIGHTED"]  # reference tests/test_filter.py:92-96


@pytest.fixture(params=BACKENDS)
def backend(request, monkeypatch):
    if request.param == "hostemu":
        emu_backend.install(monkeypatch)
    else:
        import torch

        assert torch.cuda.is_available(), "the cuda backend needs a GPU (there is no CPU fallback)"
    return request.param


@pytest.fixture()
def xr():
    mod = xr_shim.install()
    yield mod
    xr_shim.uninstall()


def scalar_data(g):
    """conftest.py:112-139 -> (GridType, data, extra_kwargs)"""
    (data,), gv = fixtures.fixture(g)
    return GridType[g], data, gv


def vector_data(g):
    """conftest.py:224-271"""
    (u, v), gv = fixtures.fixture(g)
    return GridType[g], (u, v), gv


def area_of(gv):
    for k, v in gv.items():
        if "area" in k:
            return v
    return 1


# ================================================================== tests/test_kernels.py
@pytest.mark.parametrize("g", SCALAR_GRIDS)
def test_conservation(g, backend):
    """tests/test_kernels.py:15-36: scalar Laplacians preserve the area integral."""
    grid_type, data, extra_kwargs = scalar_data(g)
    LaplacianClass = ALL_KERNELS[grid_type]
    laplacian = LaplacianClass(**extra_kwargs)
    if issubclass(LaplacianClass, AreaWeightedMixin):
        area = 1  # these act on a (transformed) regular grid with dx = dy = 1
    else:
        area = extra_kwargs.get("area", None)
        if area is None:
            area = extra_kwargs.get("tarea", 1)
    res = laplacian(data)
    np.testing.assert_allclose((area * res).sum(), 0.0, atol=1e-12)


@pytest.mark.parametrize("g", SCALAR_GRIDS + VECTOR_GRIDS)
def test_required_grid_vars(g):
    """tests/test_kernels.py:39-42, 285-288"""
    _, _, extra_kwargs = (scalar_data if g in SCALAR_GRIDS else vector_data)(g)
    assert set(required_grid_vars(GridType[g])) == set(extra_kwargs)


@pytest.mark.parametrize("g", SCALAR_GRIDS + ["MOM5U", "MOM5T"] + VECTOR_GRIDS)
def test_dimensionality(g):
    """tests/test_kernels.py:45-62, 291-299: REGULAR Laplacians are marked as nondimensional."""
    nondimensional = {"REGULAR", "REGULAR_AREA_WEIGHTED", "REGULAR_WITH_LAND", "REGULAR_WITH_LAND_AREA_WEIGHTED",
                      "TRIPOLAR_REGULAR_WITH_LAND_AREA_WEIGHTED"}
    assert ALL_KERNELS[GridType[g]].is_dimensional == (g not in nondimensional)


def test_for_large_kappas():
    """tests/test_kernels.py:68-90"""
    grid_type, _, extra_kwargs = scalar_data("IRREGULAR_WITH_LAND")
    bad_kwargs = copy.deepcopy(extra_kwargs)
    bad_kwargs["kappa_w"][99, 225] = 2.0
    with pytest.raises(ValueError, match=r"There are kappa_.*"):
        ALL_KERNELS[grid_type](**bad_kwargs)
    bad_kwargs["kappa_w"][99, 225] = 1.0
    bad_kwargs["kappa_s"][99, 225] = 2.0
    with pytest.raises(ValueError, match=r"There are kappa_.*"):
        ALL_KERNELS[grid_type](**bad_kwargs)


def test_for_kappas_not_equal_to_one():
    """tests/test_kernels.py:93-106"""
    grid_type, _, extra_kwargs = scalar_data("IRREGULAR_WITH_LAND")
    bad_kwargs = copy.deepcopy(extra_kwargs)
    bad_kwargs["kappa_w"][:, :] = 0.5
    bad_kwargs["kappa_s"][:, :] = 0.5
    with pytest.raises(ValueError, match=r"At least one place*"):
        ALL_KERNELS[grid_type](**bad_kwargs)


@pytest.mark.parametrize("direction", ["X", "Y"])
@pytest.mark.parametrize("g", IRREGULAR_GRIDS)
def test_flux(g, direction, backend):
    """tests/test_kernels.py:109-183: correct fluxes in x- and y-direction on an irregular grid; catches sign
    errors in the neighbour shifts (np.roll(dx, -1) coded as np.roll(dx, +1) and vice versa)."""
    grid_type, data, extra_kwargs = scalar_data(g)
    delta = np.zeros_like(data)
    random_yloc, random_xloc = 99, 225  # outside the mask, away from Antarctica
    delta[random_yloc, random_xloc] = 1

    test_kwargs = extra_kwargs.copy()
    for name in extra_kwargs:  # spatially uniform area, dx, dy: *isotropic* diffusion
        if not name == "wet_mask":
            test_kwargs[name] = np.ones_like(data)
    # "outlier" dx / dy just far enough from the delta that Laplacian(delta) must not feel them
    replace_data = {
        "IRREGULAR_WITH_LAND": {
            "Y": ("dxs", (random_yloc - 1, slice(None)), (random_yloc + 2, slice(None))),
            "X": ("dyw", (slice(None), random_xloc - 1), (slice(None), random_xloc + 2)),
        },
        "TRIPOLAR_POP_WITH_LAND": {
            "Y": ("dxn", (random_yloc - 2, slice(None)), (random_yloc + 1, slice(None))),
            "X": ("dye", (slice(None), random_xloc - 2), (slice(None), random_xloc + 1)),
        },
    }
    var_to_modify, slice_left, slice_right = replace_data[g][direction]
    new_data = np.ones_like(test_kwargs[var_to_modify])
    new_data[slice_left] = 1000
    new_data[slice_right] = 2000
    test_kwargs[var_to_modify] = new_data

    diffused = ALL_KERNELS[grid_type](**test_kwargs)(delta)
    np.testing.assert_allclose(diffused[random_yloc - 1, random_xloc], diffused[random_yloc + 1, random_xloc], atol=1e-12)
    np.testing.assert_allclose(diffused[random_yloc, random_xloc - 1], diffused[random_yloc, random_xloc + 1], atol=1e-12)


@pytest.mark.parametrize("g", TRIPOLAR_GRIDS)
def test_for_antarctica(g):
    """tests/test_kernels.py:189-200"""
    _, gv = fixtures.tripolar_unit_fixture(g)
    bad_kwargs = copy.deepcopy(gv)
    bad_kwargs["wet_mask"][0, 10] = 1
    with pytest.raises(AssertionError, match=r"Wet mask requires .*"):
        ALL_KERNELS[GridType[g]](**bad_kwargs)


def test_folding_of_northern_gridedge_data():
    """tests/test_kernels.py:203-221"""
    g = "TRIPOLAR_POP_WITH_LAND"
    _, gv = fixtures.tripolar_unit_fixture(g)
    bad_kwargs = copy.deepcopy(gv)
    bad_kwargs["dxn"][-1, 3] = 10
    with pytest.raises(AssertionError, match=r"Northernmost row of dxn .*"):
        ALL_KERNELS[GridType[g]](**bad_kwargs)
    bad_kwargs["dxn"][-1, 3] = 1
    bad_kwargs["dyn"][-1, 3] = 10
    with pytest.raises(AssertionError, match=r"Northernmost row of dyn .*"):
        ALL_KERNELS[GridType[g]](**bad_kwargs)


@pytest.mark.parametrize("g", TRIPOLAR_GRIDS)
def test_tripolar_exchanges(g, backend):
    """tests/test_kernels.py:224-245: exchanges across the northern boundary seam of the tripolar grid."""
    data, gv = fixtures.tripolar_unit_fixture(g)
    laplacian = ALL_KERNELS[GridType[g]](**gv)
    delta = np.zeros_like(data)
    nx = np.shape(delta)[1]
    random_loc = 10  # northern boundary, away from the edges and the pivot point in the middle
    delta[-1, random_loc] = 1
    diffused = laplacian(delta)
    # the delta diffuses isotropically across the northern boundary (regular grid data in the fixture)
    np.testing.assert_allclose(diffused[-2, random_loc], diffused[-1, nx - random_loc - 1], atol=1e-12)


@pytest.mark.parametrize("g", VECTOR_GRIDS)
def test_conservation_under_solid_body_rotation(g, backend):
    """tests/test_kernels.py:251-268: vector Laplacians are invariant under solid body rotation u = cos(lat), v = 0."""
    grid_type, _, extra_kwargs = vector_data(g)
    _, geolat_u, _, _ = fixtures.spherical_geometry()
    data_u = np.cos(geolat_u / 360 * 2 * np.pi)
    data_v = np.zeros_like(data_u)
    res_u, res_v = ALL_KERNELS[grid_type](**extra_kwargs)(data_u, data_v)
    np.testing.assert_allclose(res_u, 0.0, atol=1e-12)
    np.testing.assert_allclose(res_v, 0.0, atol=1e-12)


@pytest.mark.parametrize("g", VECTOR_GRIDS)
def test_zero_area(g, backend):
    """tests/test_kernels.py:271-282: the Laplacian must not blow up (division by zero) where areas vanish;
    on top of the reference's fixture, which has no zero areas, two patches of area_u / area_v are zeroed."""
    grid_type, (data_u, data_v), extra_kwargs = vector_data(g)
    for kwargs in (extra_kwargs, copy.deepcopy(extra_kwargs)):
        if kwargs is not extra_kwargs and g == "VECTOR_C_GRID":
            kwargs["area_u"][10:20, 30:40] = 0.0
            kwargs["area_v"][50:60, 100:110] = 0.0
        res_u, res_v = ALL_KERNELS[grid_type](**kwargs)(data_u, data_v)
        assert not np.any(np.isinf(res_u))
        assert not np.any(np.isnan(res_u))
        assert not np.any(np.isnan(res_v))


# ================================================================== tests/test_filter.py
FILTER_ARGS = dict(filter_scale=3.0, dx_min=1.0, n_steps=0, filter_shape=FilterShape.GAUSSIAN)  # test_filter.py:103-113


@pytest.mark.parametrize("g", SCALAR_GRIDS)
def test_diffusion_filter(g, backend, xr):
    """tests/test_filter.py:114-169: all diffusion-based filters (scalar Laplacians), on DataArrays."""
    grid_type, data, extra_kwargs = scalar_data(g)
    da = xr.DataArray(data, dims=["y", "x"])
    grid_vars = {name: xr.DataArray(v, dims=["y", "x"]) for name, v in extra_kwargs.items()}
    filter_args = dict(FILTER_ARGS)

    filter = Filter(grid_type=grid_type, grid_vars=grid_vars, **filter_args)
    filtered = np.asarray(filter.apply(da, dims=["y", "x"]).data)

    # conservation (xr.testing.assert_allclose: rtol 1e-5)
    area = area_of(extra_kwargs)
    np.testing.assert_allclose((data * area).sum(), (filtered * area).sum(), rtol=1e-5)

    # a scalar Laplacian cannot go through .apply_to_vector
    with pytest.raises(ValueError, match=r"Provided Laplacian *"):
        filter.apply_to_vector(da, da, dims=["y", "x"])

    # variance reduction
    assert (filtered ** 2).sum() < (data ** 2).sum()

    # an error for every missing grid variable
    for gv in grid_vars:
        grid_vars_missing = {k: v for k, v in grid_vars.items() if k != gv}
        with pytest.raises(ValueError, match=r"Provided `grid_vars` .*"):
            Filter(grid_type=grid_type, grid_vars=grid_vars_missing, **filter_args)

    bad_filter_args = copy.deepcopy(filter_args)
    bad_filter_args["transition_width"] = 1
    with pytest.raises(ValueError, match=r"Transition width .*"):
        Filter(grid_type=grid_type, grid_vars=grid_vars, **bad_filter_args)
    bad_filter_args["transition_width"] = np.pi
    bad_filter_args["ndim"] = 3
    bad_filter_args["n_steps"] = 0
    with pytest.raises(ValueError, match=r"When ndim > 2, you .*"):
        Filter(grid_type=grid_type, grid_vars=grid_vars, **bad_filter_args)
    bad_filter_args["ndim"] = 2
    bad_filter_args["n_steps"] = 3
    with pytest.warns(UserWarning, match=r"You have set n_steps .*"):
        Filter(grid_type=grid_type, grid_vars=grid_vars, **bad_filter_args)
    if g in AREA_WEIGHTED_REGULAR:
        bad_filter_args["filter_scale"] = 3
        bad_filter_args["dx_min"] = 3
        with pytest.raises(ValueError, match=r"Provided Laplacian .*"):
            Filter(grid_type=grid_type, grid_vars=grid_vars, **bad_filter_args)


def test_application_to_dataset(backend, xr):
    """tests/test_filter.py:172-218"""
    rng = np.random.default_rng(0)
    spatial = rng.normal(size=(100, 100))
    temporal = rng.normal(size=(10,))
    spatiotemporal = rng.normal(size=(10, 100, 100))
    dataset = xr.Dataset({
        "spatial": xr.DataArray(spatial.copy(), dims=["y", "x"]),
        "temporal": xr.DataArray(temporal.copy(), dims=["time"]),
        "spatiotemporal": xr.DataArray(spatiotemporal.copy(), dims=["time", "y", "x"]),
    })
    filter = Filter(filter_scale=4, dx_min=1, filter_shape=FilterShape.GAUSSIAN, grid_type=GridType.REGULAR)
    filtered_dataset = filter.apply(dataset, ["y", "x"])

    # temporal variables are unaffected: the filter only acts over space
    np.testing.assert_allclose(np.asarray(filtered_dataset["temporal"].data), temporal)
    # spatial variables change
    assert not np.allclose(np.asarray(filtered_dataset["spatial"].data), spatial)
    assert not np.allclose(np.asarray(filtered_dataset["spatiotemporal"].data), spatiotemporal)
    # spatial means of the spatiotemporal variable are unchanged
    np.testing.assert_allclose(np.asarray(filtered_dataset["spatiotemporal"].data).mean(axis=(1, 2)),
                               spatiotemporal.mean(axis=(1, 2)), rtol=1e-5, atol=1e-12)
    # the input dataset is not modified
    assert np.array_equal(np.asarray(dataset["spatial"].data), spatial)

    with pytest.warns(UserWarning, match=r".* nothing was filtered."):
        filter.apply(dataset, ["foo", "bar"])
    with pytest.warns(UserWarning, match=r".* nothing was filtered."):
        filter.apply(dataset, ["yy", "x"])


def test_nondimensional_invariance(backend, xr):
    """tests/test_filter.py:221-252: (filter_scale 4, dx_min 1) == (filter_scale 8, dx_min 2) on a REGULAR grid."""
    rng = np.random.default_rng(1)
    spatial = xr.DataArray(rng.normal(size=(100, 100)), dims=["y", "x"])
    f1 = Filter(filter_scale=4, dx_min=1, filter_shape=FilterShape.GAUSSIAN, grid_type=GridType.REGULAR)
    f2 = Filter(filter_scale=8, dx_min=2, filter_shape=FilterShape.GAUSSIAN, grid_type=GridType.REGULAR)
    a = np.asarray(f1.apply(spatial, ["y", "x"]).data)
    b = np.asarray(f2.apply(spatial, ["y", "x"]).data)
    np.testing.assert_allclose(a, b, rtol=1e-5)


@pytest.mark.parametrize("g", VECTOR_GRIDS)
def test_viscosity_filter(g, backend, xr):
    """tests/test_filter.py:256-290: all viscosity-based filters (vector Laplacians); Taper, n_steps 10."""
    filter_args = dict(filter_scale=5.0, dx_min=1.0, n_steps=10, filter_shape=FilterShape.TAPER)
    grid_type, _, extra_kwargs = vector_data(g)
    grid_vars = {name: xr.DataArray(v, dims=["y", "x"]) for name, v in extra_kwargs.items()}
    _, geolat_u, _, _ = fixtures.spherical_geometry()
    filter = Filter(grid_type=grid_type, grid_vars=grid_vars, **filter_args)

    # solid body rotation u = cos(lat), v = 0 is left alone
    data_u = np.cos(geolat_u / 360 * 2 * np.pi)
    data_v = np.zeros_like(data_u)
    da_u = xr.DataArray(data_u, dims=["y", "x"])
    da_v = xr.DataArray(data_v, dims=["y", "x"])
    filtered_u, filtered_v = filter.apply_to_vector(da_u, da_v, dims=["y", "x"])
    np.testing.assert_allclose(np.asarray(filtered_u.data), data_u, rtol=1e-5, atol=1e-12)
    np.testing.assert_allclose(np.asarray(filtered_v.data), data_v, rtol=1e-5, atol=1e-12)

    with pytest.raises(ValueError, match=r"Provided Laplacian *"):
        filter.apply(da_u, dims=["y", "x"])
    for gv in grid_vars:
        grid_vars_missing = {k: v for k, v in grid_vars.items() if k != gv}
        with pytest.raises(ValueError, match=r"Provided `grid_vars` .*"):
            Filter(grid_type=grid_type, grid_vars=grid_vars_missing, **filter_args)


def test_several_variables_in_one_call(backend):
    """Beyond the reference: a dict of variables on one grid is the array counterpart of the Dataset loop
    (reference filter.py:454-467); numpy variables of one dtype share ONE batched filter call (their batch axes are
    flattened and concatenated) and must equal the variable-by-variable results bit for bit."""
    grid_type, data, gv = scalar_data("IRREGULAR_WITH_LAND")
    rng = np.random.default_rng(11)
    variables = {"temp": data, "salt": np.stack([data * 2.0, 1.0 - data, data * data]),
                 "age": rng.random((2, 2) + data.shape)}
    flt = Filter(filter_scale=4.0, dx_min=1.0, grid_type=grid_type, grid_vars=gv)
    out = flt.apply(variables, dims=["y", "x"])
    assert set(out) == set(variables)
    for k, v in variables.items():
        single = flt.apply(v, dims=["y", "x"])
        assert out[k].shape == v.shape and np.array_equal(out[k], single, equal_nan=True), k
    assert flt.apply({}, dims=["y", "x"]) == {}
    mixed = flt.apply({"a": data, "b": data.astype(np.float32)}, dims=["y", "x"])  # mixed dtypes: one call each
    assert mixed["a"].dtype == np.float64 and np.array_equal(mixed["a"], flt.apply(data, dims=["y", "x"]), equal_nan=True)
