"""Deterministic synthetic inputs of the benchmark configurations (BASELINE.json
configs[0..4]).  Pure array generation with numpy.random.Generator(PCG64(seed)) -- no filter arithmetic, no
dependency on the oracle or on the product package.  Used by bench.py (both arms) and re-exported by
oracle/fixtures.py for the tests.  The generators follow the construction of the reference test fixtures
(/root/reference/tests/conftest.py:79-133, 180-270) at benchmark sizes."""
import numpy as np
from numpy.random import PCG64, Generator

VECTOR_C_VARS = ["wet_mask_t", "wet_mask_q", "dxT", "dyT", "dxCu", "dyCu", "dxCv", "dyCv",
                 "dxBu", "dyBu", "area_u", "area_v", "kappa_iso", "kappa_aniso"]
VECTOR_B_VARS = ["DXU", "DYU", "HUS", "HUW", "HTE", "HTN", "UAREA", "TAREA"]


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
    names = VECTOR_C_VARS if grid_type == "VECTOR_C_GRID" else VECTOR_B_VARS
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


# ------------------------------------------------------------------------------------------
# Benchmark configurations (BASELINE.json configs[0..4]).
# ``nb`` lets callers build a bounded sample (fewer batch slices) of the same workload; the
# 2-D planes never depend on nb.
# ------------------------------------------------------------------------------------------
def _batched(shape2d, nb, seed, dtype, levels=None):
    """Uniform (nb, ny, nx) field.  ``levels=(a, b)``: only slices a..b-1 of that same field (what one rank of a
    batch-sharded run needs): the PCG64 stream is advanced past the first ``a`` slices -- ``random`` draws one
    64-bit word per double -- so every rank sees exactly the values of the whole array."""
    bg = PCG64(seed)
    rng = Generator(bg)
    if nb is None:
        return rng.random(shape2d).astype(dtype, copy=False)
    a, b = (0, nb) if levels is None else levels
    if a:
        bg.advance(a * int(np.prod(shape2d)))
    return rng.random((b - a,) + tuple(shape2d)).astype(dtype, copy=False)


def cfg1(shape=(256, 512)):
    """REGULAR Gaussian, filter_scale 4, dx_min 1, fp64 (n_steps -> 5)."""
    return dict(name="cfg1", grid_type="REGULAR", fields=(uniform(shape, 100),), grid_vars={},
                filter_args=dict(filter_scale=4.0, dx_min=1.0, filter_shape="GAUSSIAN"))


def cfg2(nb=365, shape=(720, 1440), nan_land=True, levels=None):
    """REGULAR_WITH_LAND Gaussian scale 10 on a 1/4 deg fp32 field x nb daily steps (n_steps 11)."""
    f = _batched(shape, nb, 200, np.float32, levels)
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


def cfg3(nb=62, shape=(2400, 3600), dtype=np.float64, gaussian=True, nan_land=True, levels=None):
    """IRREGULAR_WITH_LAND on a POP 0.1 deg grid x nb levels.

    gaussian=True: the north-star headline (GAUSSIAN filter_scale 36, dx_min 0.9 -> n_steps 44);
    gaussian=False: BASELINE configs[2] (TAPER filter_scale 9 -> n_steps 39)."""
    f = _batched(shape, nb, 300, dtype, levels)
    gv = irregular_planes(shape, dtype)
    if nan_land:
        f[..., gv["wet_mask"] == 0] = np.nan
    fa = (dict(filter_scale=36.0, dx_min=0.9, filter_shape="GAUSSIAN") if gaussian
          else dict(filter_scale=9.0, dx_min=0.9, filter_shape="TAPER"))
    return dict(name="cfg3", grid_type="IRREGULAR_WITH_LAND", fields=(f,), grid_vars=gv, filter_args=fa)


def cfg4(nb=None, shape=(2400, 3600), levels=None):
    """TRIPOLAR_POP_WITH_LAND Gaussian, fp64, fold-symmetric dxn/dyn (n_steps 44)."""
    f = _batched(shape, nb, 400, np.float64, levels)
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


def cfgb(shape=(2160, 4320)):
    """VECTOR_B_GRID (POP B-grid viscosity operator) on the cfg5 geometry, fp64 -- not a BASELINE config; used to
    profile the one-step kernel of the fourth operator family."""
    (u, v), gv = vector_fixture("VECTOR_B_GRID", shape)
    dx_min = float(min(gv["DXU"].min(), gv["DYU"].min()))
    return dict(name="cfgb", grid_type="VECTOR_B_GRID", fields=(u, v), grid_vars=gv,
                filter_args=dict(filter_scale=20.0 * dx_min, dx_min=dx_min, filter_shape="GAUSSIAN"))
