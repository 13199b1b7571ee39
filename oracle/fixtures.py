"""Deterministic inputs: the reference test-suite fixtures and the benchmark configs.

TEST INFRASTRUCTURE ONLY.  Restates ``/root/reference/tests/conftest.py`` (same PCG64
seeds => identical arrays) so that the reference's golden zarr files apply to our
outputs, and defines the synthetic inputs of SURVEY.md section 8(d).

All arrays come from ``numpy.random.Generator(PCG64(seed))``.  The plain array generators live in
``bench_inputs.py`` (repo root) so that bench.py can build its synthetic inputs without importing the oracle.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

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


from bench_inputs import (  # noqa: E402,F401  shared array generators (no arithmetic of the path)
    cfg1, cfg2, cfg3, cfg4, cfg5, irregular_planes, land_mask, metric, metric_fold_symmetric, spherical_geometry,
    uniform, vector_fixture,
)


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


def fixture(grid_type, shape=(128, 256)):
    """Uniform access: returns (tuple_of_fields, grid_vars)."""
    if grid_type in VECTOR_FIXTURE_VARS:
        return vector_fixture(grid_type, shape)
    f, gv = scalar_fixture(grid_type, shape)
    return (f,), gv


