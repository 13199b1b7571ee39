"""Randomised check of the latitude-band decompositions (BandedFilter: one exchange per step; FusedBandedFilter:
temporal blocking, one exchange per 4-step block) over gloo on the CPU, compute through the host emulator,
against the numpy oracle on the undecomposed domain.  One rendezvous per world size, many cases per rendezvous.

    python tests/tools/fuzz_bands.py [--cases 40] [--seed 0] [--worlds 2,3,4]

Combine with an ASAN build of the emulator (GCMF_HOSTEMU_LIB, LD_PRELOAD=libasan) to catch out-of-bounds indexing
of the ghost rows.
"""
import argparse
import os
import socket
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

BAND_GRIDS = ["IRREGULAR_WITH_LAND", "TRIPOLAR_POP_WITH_LAND", "VECTOR_C_GRID", "VECTOR_B_GRID", "MOM5U", "REGULAR",
              "REGULAR_WITH_LAND_AREA_WEIGHTED", "TRIPOLAR_REGULAR_WITH_LAND_AREA_WEIGHTED", "REGULAR_WITH_LAND"]
FUSED_GRIDS = ["IRREGULAR_WITH_LAND", "REGULAR_WITH_LAND", "REGULAR_WITH_LAND_AREA_WEIGHTED", "TRIPOLAR_POP_WITH_LAND",
               "TRIPOLAR_REGULAR_WITH_LAND_AREA_WEIGHTED", "MOM5T", "REGULAR", "VECTOR_C_GRID", "VECTOR_B_GRID"]


def worker(rank, world, port, cases, seed, q):
    import torch.distributed as dist

    from gcm_filters_b200 import Filter, FilterShape, GridType
    from gcm_filters_b200.scheduler import BandedFilter, FusedBandedFilter
    from hostemu_util import emu_library
    from oracle import fixtures, np_oracle

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    bad = []
    try:
        rng = np.random.default_rng(seed)  # same stream on every rank
        for k in range(cases):
            fused = rng.random() < 0.5
            dtype = np.float32 if rng.random() < 0.3 else np.float64
            if fused:
                g = FUSED_GRIDS[rng.integers(len(FUSED_GRIDS))]
                ny = int(rng.integers(32 * world, 48 * world + 16))
                nx = int(rng.integers(128 if dtype == np.float64 else 256, 330)) // 4 * 4
                if g.startswith("VECTOR"):  # two-step blocks, two ghost rows: any band of at least four rows (emulator)
                    ny = int(rng.integers(4 * world, 30 * world))
                    nx = int(rng.integers(8, 80))
            else:
                g = BAND_GRIDS[rng.integers(len(BAND_GRIDS))]
                ny = int(rng.integers(4 * world, 30 * world))
                nx = int(rng.integers(8, 80))
            if g.startswith("TRIPOLAR") and nx % 2:
                nx += 1
            nb = int(rng.integers(1, 3))
            scale = float(rng.uniform(4.0, 9.0))
            fields, gv = fixtures.fixture(g, (ny, nx))
            fb = tuple(np.stack([f * (1 + 0.2 * b) for b in range(nb)]).astype(dtype) for f in fields)
            if "wet_mask" in gv:
                for f in fb:
                    f[:, gv["wet_mask"] == 0] = np.nan
            fa = dict(filter_scale=scale, dx_min=1.0)
            if g in fixtures.VECTOR_GRIDS:
                kx, ky = ("dxT", "dyT") if g == "VECTOR_C_GRID" else ("DXU", "DYU")
                dxm = float(min(gv[kx].min(), gv[ky].min()))
                fa = dict(filter_scale=scale * dxm, dx_min=dxm)
            gvt = {k_: v.astype(dtype) for k_, v in gv.items()}
            desc = f"#{k} world={world} {'fused' if fused else 'banded'} {g} {ny}x{nx} nb={nb} {np.dtype(dtype).name}"
            try:
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    flt = Filter(grid_type=GridType[g], grid_vars=gvt, filter_shape=FilterShape.GAUSSIAN, **fa)
                cls = FusedBandedFilter if fused else BandedFilter
                bf = cls(flt, rank, world, library=emu_library(), device="cpu")
                outs, (j0, j1) = bf.apply(*fb)
                ref = np_oracle.apply_filter(g, gv, tuple(f.astype(np.float64) for f in fb), **fa)
                ref = ref if isinstance(ref, tuple) else (ref,)
                tol = 1e-12 if dtype == np.float64 else 2e-5
                for o, r in zip(outs, ref):
                    rb = r[..., j0:j1, :]
                    if not np.array_equal(np.isnan(o), np.isnan(rb)):
                        bad.append(desc + f" rank {rank}: NaN mask mismatch")
                        break
                    ok = ~np.isnan(rb)
                    den = np.linalg.norm(rb[ok])
                    err = float(np.linalg.norm(o[ok] - rb[ok]) / den) if den > 0 else 0.0
                    if not err < tol:
                        bad.append(desc + f" rank {rank}: rel-L2 {err:.3e}")
                        break
            except Exception as exc:  # noqa: BLE001
                bad.append(desc + f" rank {rank}: {type(exc).__name__}: {exc}")
                break  # the ranks may be out of step now
        q.put((rank, bad))
    finally:
        dist.destroy_process_group()


def main():
    import torch.multiprocessing as mp

    from hostemu_util import emu_library

    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=40)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--worlds", default="2,3,4")
    args = ap.parse_args()
    emu_library()
    ctx = mp.get_context("spawn")
    failures = 0
    for world in [int(w) for w in args.worlds.split(",")]:
        with socket.socket() as s:
            s.bind(("127.0.0.1", 0))
            port = s.getsockname()[1]
        q = ctx.Queue()
        procs = [ctx.Process(target=worker, args=(r, world, port, args.cases, args.seed + world, q)) for r in range(world)]
        for p in procs:
            p.start()
        res = []
        for p in procs:
            p.join(3600)
        for p in procs:
            if p.exitcode != 0:
                print(f"world {world}: a rank exited with code {p.exitcode}")
                failures += 1
        while not q.empty():
            res.append(q.get())
        for rank, bad in sorted(res):
            for b in bad:
                print("FAIL", b)
                failures += 1
        print(f"world {world}: {args.cases} cases done, failures so far {failures}", flush=True)
    return 1 if failures else 0


if __name__ == "__main__":
    sys.exit(main())
