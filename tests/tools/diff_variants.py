"""Differential check of two host-emulator builds (e.g. default vs -DGCMF_OPT_SANSTATE=1): the same random
fused-size flux / regular problems, with NaN, +inf and -inf sprinkled over land AND ocean, must give bit-identical
results (NaN payloads aside).  Covers the inf paths that the oracle comparison leaves out (the reference itself
overflows there).

    python tests/tools/diff_variants.py A.so B.so [--cases 300] [--seed 0]
"""
import argparse
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "tools"))

import fuzz_pitch  # noqa: E402
from gcm_filters_b200 import FilterShape, GridType, _cabi  # noqa: E402
from gcm_filters_b200.filter import _compute_filter_spec, _shift_scale  # noqa: E402
from gcm_filters_b200.kernels import ALL_KERNELS  # noqa: E402
from oracle import fixtures  # noqa: E402

GRIDS = ["IRREGULAR_WITH_LAND", "MOM5U", "MOM5T", "TRIPOLAR_POP_WITH_LAND", "REGULAR_WITH_LAND", "REGULAR",
         "TRIPOLAR_REGULAR_WITH_LAND_AREA_WEIGHTED"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("libs", nargs=2)
    ap.add_argument("--cases", type=int, default=300)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    la, lb = (_cabi.Library(os.path.abspath(p)) for p in args.libs)
    rng = np.random.default_rng(args.seed)
    bad = 0
    for k in range(args.cases):
        g = GRIDS[rng.integers(len(GRIDS))]
        dtype = np.float32 if rng.random() < 0.3 else np.float64
        ny = int(rng.integers(32, 80))
        nx = int(rng.integers(128 if dtype == np.float64 else 256, 340)) // 4 * 4
        nb = int(rng.integers(1, 4))
        n_steps = int(rng.integers(3, 13))
        spb = int(rng.integers(0, 5))
        (f,), gv = fixtures.fixture(g, (ny, nx))
        fb = np.stack([f * (1 + 0.1 * b) + 0.05 * rng.standard_normal((ny, nx)) for b in range(nb)]).astype(dtype)
        if "wet_mask" in gv and rng.random() < 0.8:
            fb[:, gv["wet_mask"] == 0] = [np.nan, np.inf, -np.inf][int(rng.integers(3))]
        for junk in (np.nan, np.inf, -np.inf):  # a few anywhere
            for _ in range(int(rng.integers(0, 4))):
                fb[rng.integers(nb), rng.integers(ny), rng.integers(nx)] = junk
        if rng.random() < 0.3:  # values that overflow on their own after a few steps
            fb[rng.integers(nb), rng.integers(ny), rng.integers(nx)] = 1e307 if dtype == np.float64 else 1e37
        lap = ALL_KERNELS[GridType[g]](**{k_: v.astype(dtype) for k_, v in gv.items()})
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            spec = _compute_filter_spec(6.0, 1.0, FilterShape.GAUSSIAN, np.pi, 2, n_steps)
        c = _shift_scale(spec, lap)
        ra, _ = fuzz_pitch.run(la, lap, dtype, ny, nx, [fb], spec.p, c, rng, False, spb)
        rb, _ = fuzz_pitch.run(lb, lap, dtype, ny, nx, [fb], spec.p, c, rng, False, spb)
        if not np.array_equal(ra[0], rb[0], equal_nan=True):
            bad += 1
            d = ~((ra[0] == rb[0]) | (np.isnan(ra[0]) & np.isnan(rb[0])))
            print(f"FAIL #{k} {g} {ny}x{nx} nb={nb} {np.dtype(dtype).name} n_steps={n_steps} spb={spb}: "
                  f"{int(d.sum())} values differ, first at {tuple(np.argwhere(d)[0])}: {ra[0][d][0]} vs {rb[0][d][0]}", flush=True)
    print(f"{args.cases - bad}/{args.cases} cases identical")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
