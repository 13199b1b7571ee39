"""Generate the committed golden fixtures.  Run ONLY in the build container (needs /root/reference):

    python tests/golden/make_golden.py

Writes
  tests/golden/reference_goldens.npz   the reference's own 18 golden arrays
        (tests/test_data_kernels/*.zarr, tests/test_data_filter/*.zarr; f4), decoded from
        zarr-v2/blosc-lz4.  Keys ``kernels/<GRID>`` and ``filter/<GRID>``.
  tests/golden/ref_outputs.npz         outputs of the LIVE reference (numpy path, fp64) run
        here on the restated fixtures at 64x96 ('mid') and 37x54 ('odd'): single Laplacian and full filters for all 11 grid types
        (this is the only pin for MOM5U / MOM5T), an odd-shaped grid, NaN-on-land inputs,
        batched inputs, and FilterSpec coefficient sweeps.
Every array in these files is produced by unmodified reference code.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import fixtures, ref_loader, zarr_golden  # noqa: E402

REF_TESTS = os.path.join(ref_loader.REFERENCE_ROOT, "tests")

FILTER_CASES = {
    # name -> (filter kwargs)
    "gauss8": dict(filter_scale=8.0, dx_min=1.0, n_steps=0, filter_shape="GAUSSIAN"),
    "taper4": dict(filter_scale=4.0, dx_min=1.0, n_steps=0, filter_shape="TAPER"),
}


def odd_fixture(grid_type, shape=(37, 54)):
    """A small non-power-of-two grid (nx even so the tripolar fold is well defined)."""
    return fixtures.fixture(grid_type, shape)


def filter_args_for(grid_type, gv, fa):
    """dimensional vector fixtures use metre-scale spacings: scale the filter with dx_min."""
    fa = dict(fa)
    if grid_type in fixtures.VECTOR_GRIDS:
        key = "dxT" if grid_type == "VECTOR_C_GRID" else "DXU"
        dxm = float(min(gv[key].min(), gv["dyT" if grid_type == "VECTOR_C_GRID" else "DYU"].min()))
        fa["dx_min"] = dxm
        fa["filter_scale"] = fa["filter_scale"] * dxm
    return fa


def main():
    ref_loader.load()
    out = {}
    for kind, sub in (("kernels", "test_data_kernels"), ("filter", "test_data_filter")):
        for g in fixtures.GOLDEN_SCALAR_GRIDS + fixtures.VECTOR_GRIDS:
            out[f"{kind}/{g}"] = zarr_golden.read_zarr_array(os.path.join(REF_TESTS, sub, f"{g}.zarr"))
    np.savez_compressed(os.path.join(HERE, "reference_goldens.npz"), **out)
    print("reference_goldens.npz:", len(out), "arrays")

    live = {}
    for g in fixtures.SCALAR_GRIDS + fixtures.VECTOR_GRIDS:
        for tag, (fields, gv) in (("mid", fixtures.fixture(g, (64, 96))), ("odd", odd_fixture(g))):
            lap = ref_loader.ref_laplacian(g, gv, *fields)
            live[f"lap/{tag}/{g}"] = np.stack(lap) if isinstance(lap, tuple) else lap
            for cname, fa in FILTER_CASES.items():
                if tag == "odd" and cname != "gauss8":
                    continue
                res, _ = ref_loader.ref_filter(g, gv, fields, **filter_args_for(g, gv, fa))
                live[f"filter/{tag}/{cname}/{g}"] = np.stack(res) if isinstance(res, tuple) else res
    # NaN on land + batch dims for the land-aware scalar grids
    for g in ("REGULAR_WITH_LAND", "IRREGULAR_WITH_LAND", "TRIPOLAR_POP_WITH_LAND",
              "TRIPOLAR_REGULAR_WITH_LAND_AREA_WEIGHTED", "MOM5T"):
        (f,), gv = fixtures.fixture(g, (48, 64))
        fb = np.stack([f, f[::-1].copy(), f * f])
        fb[:, gv["wet_mask"] == 0] = np.nan
        res, _ = ref_loader.ref_filter(g, gv, (fb,), **FILTER_CASES["gauss8"])
        live[f"filter/nanbatch/gauss8/{g}"] = res
    # FilterSpec sweeps (filter.py:99-151)
    gf = ref_loader.load()
    from gcm_filters.filter import _compute_filter_spec, _compute_n_steps_default
    sweep = []
    for shape in ("GAUSSIAN", "TAPER"):
        for ndim in (1, 2):
            for ratio in (1.5, 4.0, 10.0, 40.0):
                for tw in (np.pi, 2.0):
                    n = int(_compute_n_steps_default(ndim, gf.FilterShape[shape], ratio * 0.9, 0.9, tw))
                    spec = _compute_filter_spec(ratio * 0.9, 0.9, gf.FilterShape[shape], tw, ndim, n)
                    key = f"spec/{shape}/{ndim}/{ratio}/{tw:.4f}"
                    live[key] = np.concatenate(([spec.n_steps, spec.s_max, spec.dx_min_sq], spec.p))
                    sweep.append(key)
    np.savez_compressed(os.path.join(HERE, "ref_outputs.npz"), **live)
    print("ref_outputs.npz:", len(live), "arrays")


if __name__ == "__main__":
    main()
