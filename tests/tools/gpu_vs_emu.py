"""Development check (needs a GPU and g++): is the CUDA build bit-identical to the host emulator of the same sources?

    python tests/tools/gpu_vs_emu.py

Runs one Laplacian and one filter per operator family and dtype through gcm_filters_b200 (CUDA) and through
tests/emu_backend.py (the emulator), and reports the largest difference.  Everything on the path is IEEE +, -, *, /
and explicit fma with contraction off on both sides, so the expectation is 0 everywhere; a non-zero entry points at
a code path where nvcc and g++ disagree (contraction, division, a different evaluation order).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import emu_backend  # noqa: E402
from gcm_filters_b200 import Filter, GridType  # noqa: E402
from gcm_filters_b200.filter import _shift_scale  # noqa: E402
from gcm_filters_b200.kernels import ALL_KERNELS  # noqa: E402
from oracle import fixtures  # noqa: E402


def main():
    worst = 0.0
    for g in fixtures.SCALAR_GRIDS + fixtures.VECTOR_GRIDS:
        for dtype in (np.float64, np.float32):
            fields, gv = fixtures.fixture(g, (96, 160))
            fields = tuple(f.astype(dtype) for f in fields)
            gvt = {k: v.astype(dtype) for k, v in gv.items()}
            fa = dict(filter_scale=8.0, dx_min=1.0)
            if g in fixtures.VECTOR_GRIDS:
                kx, ky = ("dxT", "dyT") if g == "VECTOR_C_GRID" else ("DXU", "DYU")
                dxm = float(min(gv[kx].min(), gv[ky].min()))
                fa = dict(filter_scale=8.0 * dxm, dx_min=dxm)
            flt = Filter(grid_type=GridType[g], grid_vars=gvt, **fa)
            lap = ALL_KERNELS[GridType[g]](**gvt)
            c = _shift_scale(flt.filter_spec, flt.laplacian)
            if len(fields) == 2:
                gpu = list(lap(*fields)) + list(flt.apply_to_vector(*fields, dims=["y", "x"]))
            else:
                gpu = [lap(fields[0]), flt.apply(fields[0], dims=["y", "x"])]
            emu = list(emu_backend.run_laplacian(lap, fields)) + list(
                emu_backend.run_filter(flt.laplacian, flt.filter_spec.p, c, fields))
            d = max(float(np.max(np.abs(np.nan_to_num(np.asarray(a, dtype=np.float64)) -
                                        np.nan_to_num(np.asarray(b, dtype=np.float64))))) for a, b in zip(gpu, emu))
            same_nan = all(np.array_equal(np.isnan(a), np.isnan(b)) for a, b in zip(gpu, emu))
            worst = max(worst, d)
            print(f"{g:45s} {np.dtype(dtype).name:8s} max|gpu-emu| = {d:.3e}  nan masks equal: {same_nan}", flush=True)
    print("bit-identical" if worst == 0.0 else f"largest difference {worst:.3e}")


if __name__ == "__main__":
    main()
