"""Deterministic inputs: the reference test-suite fixtures and the benchmark configs.

TEST INFRASTRUCTURE ONLY.  Restates ``/root/reference/tests/conftest.py`` (same PCG64
seeds => identical arrays) so that the reference's golden zarr files apply to our
outputs, and defines the synthetic inputs of SURVEY.md section 8(d).

All arrays come from ``numpy.random.Generator(PCG64(seed))``.
"""
import numpy as np
from numpy.random import PCG64, Generator

# positional grid-var order used by the reference *fixtures* (conftest.py:12-31); the seed of
# a metric array is its index in this list (conftest.py:118-124).
SCALAR_FIXTURE_VARS = {
    "REGULAR": [],
    "REGULAR_AREA_WEIGHTED": ["area"],
    "REGULAR_WITH_LAND": ["wet_mask"],
    "REGULAR_WITH_LAND_AREA_WEIGHTED": ["wet_mask", "area"],
    "IRREGULAR_WITH_LAND": ["wet_mask", "dxw", "dyw", "dxs", "dys", "area", "kappa_w", "kappa_s"],
    "MOM5U": ["wet_mask", "dxt", "dyt", "dxu", "dyu", "area_u"],
    "MOM5T": ["wet_mask", "dxt", "dyt", "dxu", "dyu", "area_t"],
    "TRIPOLAR_REGULAR_WITH_LAND_AREA_WEIGHTED": ["wet_mask", "area"],
    "TRIPOLAR_POP_WITH_LAND": ["wet_mask", "dxe", "dye", "dxn", "dyn", "tarea"],
}
VECTOR_FIXTURE_VARS = {
    "VECTOR_C_GRID": ["wet_mask_t", "wet_mask_q", "dxT", "dyT", "dxCu", "dyCu", "dxCv", "dyCv",
                      "dxBu", "dyBu", "area_u", "area_v", "kappa_iso", "kappa_aniso"],
    "VECTOR_B_GRID": ["DXU", "DYU", "HUS", "HUW", "HTE", "HTN", "UAREA", "TAREA"],
}
SCALAR_GRIDS = list(SCALAR_FIXTURE_VARS)
VECTOR_GRIDS = list(VECTOR_FIXTURE_VARS)
# grid types for which the reference ships goldens (conftest.py:62-70): no MOM5U/T
GOLDEN_SCALAR_GRIDS = [g for g in SCALAR_GRIDS if not g.startswith("MOM5")]


def uniform(shape, seed):
    """conftest.py:79-81"""
    return Generator(PCG64(seed)).random(shape)


def land_mask(shape, dtype=np.float64):
    """conftest.py:84-89: row 0 (Antarctica) and the SW quadrant are land."""
    ny, nx = shape
    m = np.ones(shape, dtype=dtype)
    m[0, :] = 0
    m[: ny // 2, : nx // 2] = 0
    return m


def metric(shape, seed):
    """conftest.py:92-97: positive, mean 1, +-10 %."""
    return 0.9 + 0.2 * Generator(PCG64(seed)).random(shape)


def metric_fold_symmetric(shape, seed):
    """conftest.py:100-109: as ``metric`` but the last row mirrors onto itself."""
    g = metric(shape, seed)
    nx = shape[-1]
    g[-1, nx // 2:] = g[-1, : nx // 2][::-1]
    return g


def scalar_fixture(grid_type, shape=(128, 256), data_seed=100):
    """conftest.py:112-133.  Returns (field, grid_vars dict)."""
    names = SCALAR_FIXTURE_VARS[grid_type]
    field = uniform(shape, data_seed)
    gv = {}
    seed = -1
    for seed, name in enumerate(names):
        if name == "wet_mask":
            gv[name] = land_mask(shape)
        elif "kappa" in name:
            gv[name] = np.ones(shape)
        else:
            gv[name] = metric(shape, seed)
    if grid_type == "TRIPOLAR_POP_WITH_LAND":
        for name in names:
            if name in ("dxn", "dyn"):
                seed += 1
                gv[name] = metric_fold_symmetric(shape, seed)
    return field, gv


def tripolar_unit_fixture(grid_type, shape=(128, 256)):
    """conftest.py:146-162: data PCG64(30), all metrics one."""
    field = uniform(shape, 30)
    gv = {}
    for name in SCALAR_FIXTURE_VARS[grid_type]:
        gv[name] = land_mask(shape) if name == "wet_mask" else np.ones(shape)
    return field, gv


def spherical_geometry(ny=128, nx=256, lat=(-70.0, 70.0), lon=(0.0, 60.0)):
    """conftest.py:180-213: lat/lon of the u- and v-points of a spherical C-grid."""
    la0, la1 = lat
    lo0, lo1 = lon
    latCu = np.linspace(la0 + 0.5 * (la1 - la0) / ny, la1 - 0.5 * (la1 - la0) / ny, ny)
    latCv = np.linspace(la0 + (la1 - la0) / ny, la1, ny)
    lonCu = np.linspace(lo0 + (lo1 - lo0) / nx, lo1, nx)
    lonCv = np.linspace(lo0 + 0.5 * (lo1 - lo0) / nx, lo1 - 0.5 * (lo1 - lo0) / nx, nx)
    geolonCu, geolatCu = np.meshgrid(lonCu, latCu)
    geolonCv, geolatCv = np.meshgrid(lonCv, latCv)
    return geolonCu, geolatCu, geolonCv, geolatCv


def vector_fixture(grid_type, shape=(128, 256)):
    """conftest.py:216-270.  Returns ((u, v), grid_vars)."""
    ny, nx = shape
    _, geolatCu, _, geolatCv = spherical_geometry(ny, nx)
    names = VECTOR_FIXTURE_VARS[grid_type]
    R = 6378000
    dx_u = R * np.cos(geolatCu / 360 * 2 * np.pi)
    dx_v = R * np.cos(geolatCv / 360 * 2 * np.pi)
    dy = np.max(dx_u) * np.ones((ny, nx))
    gv = {}
    for name in names:
        if name in ("dxCu", "dxT", "HUS", "HTE"):
            gv[name] = dx_u.copy()
        elif name in ("dxCv", "dxBu", "DXU", "HUW", "HTN"):
            gv[name] = dx_v.copy()
        elif name in ("dyCu", "dyCv", "dyBu", "dyT", "DYU"):
            gv[name] = dy.copy()
    for name in names:
        if name == "area_u":
            gv[name] = gv["dxCu"] * gv["dyCu"]
        elif name == "area_v":
            gv[name] = gv["dxCv"] * gv["dyCv"]
        elif name == "UAREA":
            gv[name] = gv["DXU"] * gv["DYU"]
        elif name == "TAREA":
            gv[name] = gv["HTE"] * gv["DYU"]
        elif name in ("kappa_iso", "kappa_aniso"):
            gv[name] = np.ones((ny, nx))
    island = np.ones((ny, nx))
    island[: ny // 2, : nx // 2] = 0
    for name in names:
        if name in ("wet_mask_t", "wet_mask_q"):
            gv[name] = island.copy()
    u = uniform((ny, nx), 42)
    v = uniform((ny, nx), 43)
    return (u, v), gv


def fixture(grid_type, shape=(128, 256)):
    """Uniform access: returns (tuple_of_fields, grid_vars)."""
    if grid_type in VECTOR_FIXTURE_VARS:
        return vector_fixture(grid_type, shape)
    f, gv = scalar_fixture(grid_type, shape)
    return (f,), gv


# ------------------------------------------------------------------------------------------
# Benchmark configurations (BASELINE.json configs[0..4]).
# ``nb`` lets callers build a bounded sample (fewer batch slices) of the same workload; the
# 2-D planes never depend on nb.
# ------------------------------------------------------------------------------------------
def _batched(shape2d, nb, seed, dtype):
    rng = Generator(PCG64(seed))
    if nb is None:
        return rng.random(shape2d).astype(dtype, copy=False)
    return rng.random((nb,) + tuple(shape2d)).astype(dtype, copy=False)


def cfg1(shape=(256, 512)):
    """REGULAR Gaussian, filter_scale 4, dx_min 1, fp64 (n_steps -> 5)."""
    return dict(name="cfg1", grid_type="REGULAR", fields=(uniform(shape, 100),), grid_vars={},
                filter_args=dict(filter_scale=4.0, dx_min=1.0, filter_shape="GAUSSIAN"))


def cfg2(nb=365, shape=(720, 1440), nan_land=True):
    """REGULAR_WITH_LAND Gaussian scale 10 on a 1/4 deg fp32 field x nb daily steps (n_steps 11)."""
    f = _batched(shape, nb, 200, np.float32)
    m = land_mask(shape, np.float32)
    if nan_land:
        f[..., m == 0] = np.nan
    return dict(name="cfg2", grid_type="REGULAR_WITH_LAND", fields=(f,), grid_vars={"wet_mask": m},
                filter_args=dict(filter_scale=10.0, dx_min=1.0, filter_shape="GAUSSIAN"))


def irregular_planes(shape, dtype=np.float64):
    ny, nx = shape
    gv = {"wet_mask": land_mask(shape, dtype)}
    for seed, name in enumerate(["dxw", "dyw", "dxs", "dys", "area"], start=1):
        gv[name] = metric(shape, seed).astype(dtype, copy=False)
    for seed, name in ((6, "kappa_w"), (7, "kappa_s")):
        k = 0.5 + 0.5 * Generator(PCG64(seed)).random(shape)
        k[ny - 1, nx - 1] = 1.0  # kernels.py:274-281: some kappa must equal 1
        gv[name] = k.astype(dtype, copy=False)
    return gv


def cfg3(nb=62, shape=(2400, 3600), dtype=np.float64, gaussian=True, nan_land=True):
    """IRREGULAR_WITH_LAND on a POP 0.1 deg grid x nb levels.

    gaussian=True: the north-star headline (GAUSSIAN filter_scale 36, dx_min 0.9 -> n_steps 44);
    gaussian=False: BASELINE configs[2] (TAPER filter_scale 9 -> n_steps 39)."""
    f = _batched(shape, nb, 300, dtype)
    gv = irregular_planes(shape, dtype)
    if nan_land:
        f[..., gv["wet_mask"] == 0] = np.nan
    fa = (dict(filter_scale=36.0, dx_min=0.9, filter_shape="GAUSSIAN") if gaussian
          else dict(filter_scale=9.0, dx_min=0.9, filter_shape="TAPER"))
    return dict(name="cfg3", grid_type="IRREGULAR_WITH_LAND", fields=(f,), grid_vars=gv, filter_args=fa)


def cfg4(nb=None, shape=(2400, 3600)):
    """TRIPOLAR_POP_WITH_LAND Gaussian, fp64, fold-symmetric dxn/dyn (n_steps 44)."""
    f = _batched(shape, nb, 400, np.float64)
    gv = {"wet_mask": land_mask(shape)}
    gv["dxe"] = metric(shape, 1)
    gv["dye"] = metric(shape, 2)
    gv["dxn"] = metric_fold_symmetric(shape, 6)
    gv["dyn"] = metric_fold_symmetric(shape, 7)
    gv["tarea"] = metric(shape, 5)
    return dict(name="cfg4", grid_type="TRIPOLAR_POP_WITH_LAND", fields=(f,), grid_vars=gv,
                filter_args=dict(filter_scale=36.0, dx_min=0.9, filter_shape="GAUSSIAN"))


def cfg5(shape=(2160, 4320)):
    """VECTOR_C_GRID on a MOM6 1/12 deg-like spherical C-grid, fp64 (n_steps 22)."""
    (u, v), gv = vector_fixture("VECTOR_C_GRID", shape)
    wet = gv["wet_mask_t"] > 0
    dx_min = float(min(gv["dxT"][wet].min(), gv["dyT"][wet].min()))
    return dict(name="cfg5", grid_type="VECTOR_C_GRID", fields=(u, v), grid_vars=gv,
                filter_args=dict(filter_scale=20.0 * dx_min, dx_min=dx_min, filter_shape="GAUSSIAN"))
