"""Randomised check of the CUDA source (through the host emulator) against the numpy oracle: grid type, shape (odd
and vector-multiple widths, below and above the fused tile size), batch count, dtype, land pattern, NaNs on land,
number of steps and the steps-per-block cap are drawn at random.  CPU only.

    python tests/tools/fuzz_hostemu.py [--cases 200] [--seed 0]
"""
import argparse
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import rel_l2  # noqa: E402
from gcm_filters_b200 import FilterShape, GridType  # noqa: E402
from gcm_filters_b200.filter import _compute_filter_spec, _shift_scale  # noqa: E402
from gcm_filters_b200.kernels import ALL_KERNELS  # noqa: E402
from hostemu_util import EmuPlan, emu_set_steps_per_block  # noqa: E402
from oracle import fixtures, np_oracle  # noqa: E402

GRIDS = fixtures.SCALAR_GRIDS + fixtures.VECTOR_GRIDS


def one_case(rng, k):
    g = GRIDS[rng.integers(len(GRIDS))]
    big = rng.random() < 0.5  # large enough for the fused kernel (tile 32 x 128 fp64 / 32 x 256 fp32)
    dtype = np.float32 if rng.random() < 0.35 else np.float64
    if big:
        ny = int(rng.integers(32, 90))
        nx = int(rng.integers(128 if dtype == np.float64 else 256, 400))
        if rng.random() < 0.8:
            nx -= nx % 4
    else:
        ny, nx = int(rng.integers(6, 40)), int(rng.integers(6, 90))
    if g.startswith("TRIPOLAR") and nx % 2:
        nx += 1  # the fold pairs column i with nx-1-i
    nb = int(rng.integers(1, 4))
    n_steps = int(rng.integers(3, 12))
    spb = int(rng.integers(0, 5))
    fields, gv = fixtures.fixture(g, (ny, nx))
    gv = {k_: v.copy() for k_, v in gv.items()}
    masks = [k_ for k_ in gv if "mask" in k_]
    if masks and rng.random() < 0.6:  # random land on top of the fixture's pattern
        land = rng.random((ny, nx)) < rng.uniform(0.0, 0.4)
        for m in masks:
            gv[m] = gv[m] * (~land)
    if "kappa_w" in gv:
        gv["kappa_w"] = 0.5 + 0.5 * rng.random((ny, nx))
        gv["kappa_s"] = 0.5 + 0.5 * rng.random((ny, nx))
        gv["kappa_w"][ny // 2, nx // 2] = 1.0
    fb = [np.stack([f * (1 + 0.1 * b) + 0.05 * rng.standard_normal((ny, nx)) for b in range(nb)]) for f in fields]
    if "wet_mask" in gv and rng.random() < 0.7:
        junk = [np.nan, np.inf, -np.inf][int(rng.integers(3))]  # nan_to_num: NaN -> 0, +-inf -> +-largest finite
        if "dxw" in gv or "dxe" in gv or "dxt" in gv:
            # flux-form operators difference nan_to_num(f) across land-sea faces before masking the face: with +-inf
            # on land the reference overflows ((1.8e308 - x) / 0.9 = inf, inf * 0 = NaN) and poisons the neighbouring
            # ocean cells; the precombined face coefficient here is an exact 0.  Undefined input: NaN only.
            junk = np.nan
        for f in fb:
            f[:, gv["wet_mask"] == 0] = junk
    elif rng.random() < 0.3:  # a few NaNs in the open ocean: they spread (or are zeroed) exactly as in the reference
        for f in fb:
            f[rng.integers(nb), rng.integers(ny), rng.integers(nx)] = np.nan
    if nb > 1 and not g.startswith("MOM5") and rng.random() < 0.25:
        # batched grid variables: one plane per batch slice, e.g. a wet mask per depth level
        for k_ in list(gv):
            if "mask" in k_:
                lvl = np.stack([gv[k_] * (rng.random((ny, nx)) > 0.1 * b) for b in range(nb)])
                if g.startswith("TRIPOLAR"):
                    lvl[:, 0, :] = 0
                gv[k_] = lvl
            elif "kappa" not in k_:
                gv[k_] = np.stack([gv[k_] * (1.0 + 0.05 * b) for b in range(nb)])
    dxm = 1.0
    if g in fixtures.VECTOR_GRIDS:
        kx, ky = ("dxT", "dyT") if g == "VECTOR_C_GRID" else ("DXU", "DYU")
        dxm = float(min(gv[kx].min(), gv[ky].min()))
    gvt = {k_: v.astype(dtype) for k_, v in gv.items()}
    lap = ALL_KERNELS[GridType[g]](**gvt)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        spec = _compute_filter_spec(6.0 * dxm, dxm, FilterShape.GAUSSIAN, np.pi, 2, n_steps)
    c = _shift_scale(spec, lap)
    plan = EmuPlan(lap, dtype, ny, nx)
    emu_set_steps_per_block(plan, spb)
    fin = tuple(f.astype(dtype) for f in fb)
    got = plan.filter(fin, spec.p, c)
    lap_got = plan.laplacian(fin)
    op = np_oracle.make_operator(g, gv)
    ref = np_oracle.run_recurrence(op, np_oracle.FilterSpec(*spec), tuple(f.astype(dtype).astype(np.float64) for f in fb))
    ref = ref if isinstance(ref, tuple) else (ref,)
    tol = 1e-12 if dtype == np.float64 else 2e-5
    desc = f"#{k} {g} {ny}x{nx} nb={nb} {np.dtype(dtype).name} n_steps={n_steps} spb={spb}"
    for a, b in zip(got, ref):
        if not np.array_equal(np.isnan(a), np.isnan(b)):
            return desc + " NaN MASK MISMATCH"
        err = rel_l2(a, b)
        if not err < tol:
            return desc + f" rel-L2 {err:.3e}"
    lref = np_oracle.laplacian(g, gv, *[f.astype(dtype).astype(np.float64) for f in fb])
    lref = lref if isinstance(lref, tuple) else (lref,)
    for a, b in zip(lap_got, lref):
        if not np.array_equal(np.isnan(a), np.isnan(b)):
            return desc + " laplacian NaN MASK MISMATCH"
        err = rel_l2(a, b)
        if not err < (1e-13 if dtype == np.float64 else 1e-4):
            return desc + f" laplacian rel-L2 {err:.3e}"
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=200)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    rng = np.random.default_rng(args.seed)
    bad = 0
    for k in range(args.cases):
        try:
            msg = one_case(rng, k)
        except Exception as exc:  # noqa: BLE001
            msg = f"#{k} raised {type(exc).__name__}: {exc}"
        if msg:
            bad += 1
            print("FAIL", msg, flush=True)
    print(f"{args.cases - bad}/{args.cases} cases ok")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
