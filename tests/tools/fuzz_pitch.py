"""Randomised check of the C ABI's strided fields and planes (row pitch > nx, batch stride > ny*pitch, unaligned base
pointers -> scalar fallback of the vector kernels): results must be bit-identical to the contiguous call.  CPU only
(host emulator); combine with the ASAN build via GCMF_HOSTEMU_LIB.

    python tests/tools/fuzz_pitch.py [--cases 200] [--seed 0]
"""
import argparse
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from gcm_filters_b200 import FilterShape, GridType, _cabi  # noqa: E402
from gcm_filters_b200.filter import _compute_filter_spec, _shift_scale  # noqa: E402
from gcm_filters_b200.kernels import ALL_KERNELS  # noqa: E402
from hostemu_util import _DT, emu_library  # noqa: E402
from oracle import fixtures  # noqa: E402

GRIDS = fixtures.SCALAR_GRIDS + fixtures.VECTOR_GRIDS


def padded(rng, a, dtype, keep):
    """Copy of (nb, ny, nx) array `a` inside a larger buffer: returns (ptr, pitch, bstride, view)."""
    nb, ny, nx = a.shape
    es = np.dtype(dtype).itemsize
    pitch = nx + int(rng.integers(0, 9))
    bstride = ny * pitch + int(rng.integers(0, 17))
    off = int(rng.integers(0, 5))  # elements: 0 keeps the 16-byte alignment, others force the scalar path
    buf = np.full(off + nb * bstride + 64, 7777.0 if np.dtype(dtype).kind == "f" else 7, dtype=dtype)
    keep.append(buf)
    for b in range(nb):
        for j in range(ny):
            s = off + b * bstride + j * pitch
            buf[s:s + nx] = a[b, j]
    return buf.ctypes.data + off * es, pitch, bstride, (buf, off)


def unpad(view, nb, ny, nx, pitch, bstride):
    buf, off = view
    out = np.empty((nb, ny, nx), dtype=buf.dtype)
    for b in range(nb):
        for j in range(ny):
            s = off + b * bstride + j * pitch
            out[b, j] = buf[s:s + nx]
    return out


def run(lib, lap, dtype, ny, nx, fields, p, c, rng, pad, spb, what="filter"):
    spec = lap._planes
    h = lib.plan_create(spec.op, _DT[np.dtype(dtype)], ny, nx, spec.flags, 0)
    keep = []
    for slot, pl in enumerate(spec.planes):
        is_mask = slot == 0 and spec.op == _cabi.OP_REGULAR5
        src = spec.mask if is_mask else pl
        if src is None:
            continue
        a = np.ascontiguousarray(src, dtype=np.uint8 if is_mask else dtype).reshape((-1, ny, nx))
        if pad:
            prng = np.random.default_rng(12345) if spec.op == _cabi.OP_VECTOR_C else rng  # C-grid planes share one pitch
            ptr, pitch, bstride, _ = padded(prng, a, a.dtype, keep)
        else:
            keep.append(a)
            ptr, pitch, bstride = a.ctypes.data, nx, ny * nx
        lib.plan_set_plane(h, slot, ptr, pitch, bstride, a.shape[0])
    lib.plan_set_filter(h, p, c)
    lib.set_steps_per_block(h, spb)
    nb = fields[0].shape[0]
    fin, fout, views = [], [], []
    share = spec.op == _cabi.OP_VECTOR_C  # the C-grid kernels want one row pitch for u and v (EINVAL otherwise)
    fseed = int(rng.integers(1 << 30))
    for f in fields:
        if pad:
            ptr, pitch, bstride, _ = padded(np.random.default_rng(fseed) if share else rng, f, dtype, keep)
            fin.append((ptr, pitch, bstride))
            optr, opitch, obstride, view = padded(rng, np.zeros_like(f), dtype, keep)
            fout.append((optr, opitch, obstride))
            views.append((view, opitch, obstride))
        else:
            keep.append(f)
            o = np.zeros_like(f)
            keep.append(o)
            fin.append((f.ctypes.data, nx, ny * nx))
            fout.append((o.ctypes.data, nx, ny * nx))
            views.append(o)
    nbytes = lib.workspace_bytes(h, nb)
    raw = np.zeros(nbytes + 256, dtype=np.uint8)
    off = (-raw.ctypes.data) % 256
    if what == "filter":
        lib.filter(h, nb, fin, fout, raw.ctypes.data + off, nbytes)
    elif what == "laplacian":
        lib.laplacian(h, nb, fin, fout)
    elif what == "prepare":
        lib.prepare(h, nb, fin, fout)
    else:
        lib.finalize(h, nb, fin, fout)
    lib.plan_destroy(h)
    if pad:
        for (buf, off), pi, bs in views:  # nothing outside the (nb, ny, nx) elements of an output may be written
            legit = np.zeros(buf.shape, dtype=bool)
            for b in range(nb):
                for j in range(ny):
                    legit[off + b * bs + j * pi: off + b * bs + j * pi + nx] = True
            if not np.all(buf[~legit] == 7777.0):
                raise AssertionError("padding of an output field was overwritten")
        return [unpad(v, nb, ny, nx, pi, bs) for v, pi, bs in views], keep
    return views, keep


def one_case(rng, k):
    g = GRIDS[rng.integers(len(GRIDS))]
    dtype = np.float32 if rng.random() < 0.4 else np.float64
    big = rng.random() < 0.4
    ny = int(rng.integers(32, 70)) if big else int(rng.integers(6, 40))
    nx = (int(rng.integers(128 if dtype == np.float64 else 256, 330)) // 4 * 4) if big else int(rng.integers(6, 90))
    if g.startswith("TRIPOLAR") and nx % 2:
        nx += 1
    nb = int(rng.integers(1, 4))
    n_steps = int(rng.integers(3, 10))
    spb = int(rng.integers(0, 5))
    fields, gv = fixtures.fixture(g, (ny, nx))
    fb = [np.ascontiguousarray(np.stack([f * (1 + 0.1 * b) for b in range(nb)]).astype(dtype)) for f in fields]
    if "wet_mask" in gv:
        for f in fb:
            f[:, gv["wet_mask"] == 0] = np.nan
    dxm = 1.0
    if g in fixtures.VECTOR_GRIDS:
        kx, ky = ("dxT", "dyT") if g == "VECTOR_C_GRID" else ("DXU", "DYU")
        dxm = float(min(gv[kx].min(), gv[ky].min()))
    lap = ALL_KERNELS[GridType[g]](**{k_: v.astype(dtype) for k_, v in gv.items()})
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        spec = _compute_filter_spec(6.0 * dxm, dxm, FilterShape.GAUSSIAN, np.pi, 2, n_steps)
    c = _shift_scale(spec, lap)
    lib = emu_library()
    ref, _k1 = run(lib, lap, dtype, ny, nx, fb, spec.p, c, rng, False, spb)
    got, keep = run(lib, lap, dtype, ny, nx, fb, spec.p, c, rng, True, spb)
    desc = f"#{k} {g} {ny}x{nx} nb={nb} {np.dtype(dtype).name} n_steps={n_steps} spb={spb}"
    # the other entry points of the boundary: one Laplacian, and the area prepare / finalize of the operator protocol
    for what in ("laplacian",) + (("prepare", "finalize") if len(fb) == 1 else ()):
        r1, _ = run(lib, lap, dtype, ny, nx, fb, spec.p, c, rng, False, spb, what)
        r2, _ = run(lib, lap, dtype, ny, nx, fb, spec.p, c, rng, True, spb, what)
        for a, b in zip(r2, r1):
            if not np.array_equal(a, b, equal_nan=True):
                return desc + f" strided {what} differs from the contiguous one"
    for a, b in zip(got, ref):
        if g.startswith("TRIPOLAR"):
            # strided planes / fields may take the one-step kernels where the contiguous call is fused; across a
            # tripolar fold the two agree to rounding only (mirrored cells sum their fluxes in the opposite order)
            if not np.array_equal(np.isnan(a), np.isnan(b)):
                return desc + " NaN masks differ"
            ok = ~np.isnan(b)
            err = np.linalg.norm(a[ok].astype(np.float64) - b[ok]) / max(np.linalg.norm(b[ok].astype(np.float64)), 1e-300)
            if not err < (1e-14 if dtype == np.float64 else 1e-6):
                return desc + f" strided result differs from the contiguous one by {err:.2e}"
        elif not np.array_equal(a, b, equal_nan=True):
            return desc + " strided result differs from the contiguous one"
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
