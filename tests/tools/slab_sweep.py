"""Short-slab behaviour of the fused flux kernel (development tool, needs a GPU).

    python tests/tools/slab_sweep.py [--lib path/to/libgcmf.so]

Times gcmf_filter on the cfg3-shaped problem (2400 x 3600 fp64, 44 steps = 11 fused launches) for several batch
sizes nb and level-slab lengths (GCMF_FUSED_LEVELS_PER_CTA): what one GPU of an N-GPU batch split sees (62 levels over
8 GPUs = 8 levels each), and how much of a launch is per-CTA prologue (barrier set-up, coefficient tiles, the first
un-overlapped state tiles) against per-level work.  One JSON line per configuration.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from gcm_filters_b200 import _cabi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", default=None)
    ap.add_argument("--ny", type=int, default=2400)
    ap.add_argument("--nx", type=int, default=3600)
    ap.add_argument("--steps", type=int, default=44)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--nbs", default="1,2,4,7,8,16,31,62")
    ap.add_argument("--lpcs", default="0", help="comma list of forced slab lengths per nb (0 = the library's own choice)")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    lib = _cabi.Library(os.path.abspath(args.lib)) if args.lib else _cabi.get_library()
    g = torch.Generator(device=dev).manual_seed(1)
    ny, nx = args.ny, args.nx
    tdt = torch.float64
    nbmax = max(int(v) for v in args.nbs.split(","))
    field = torch.rand((nbmax, ny, nx), generator=g, device=dev, dtype=tdt)
    field[:, : ny // 2, : nx // 2] = float("nan")
    wet = torch.ones((ny, nx), device=dev, dtype=tdt)
    wet[: ny // 2, : nx // 2] = 0
    ce = (0.9 + 0.2 * torch.rand((ny, nx), generator=g, device=dev, dtype=tdt)) * wet * torch.roll(wet, -1, 1)
    cn = (0.9 + 0.2 * torch.rand((ny, nx), generator=g, device=dev, dtype=tdt)) * wet * torch.roll(wet, -1, 0)
    ra = 1.0 / (0.9 + 0.2 * torch.rand((ny, nx), generator=g, device=dev, dtype=tdt))
    p = [1.0 / (i + 2) * (-1) ** i for i in range(args.steps + 1)]
    out = torch.empty_like(field)
    stream = torch.cuda.current_stream(dev)
    h = lib.plan_create(_cabi.OP_FLUX, _cabi.GCMF_F64, ny, nx, _cabi.FLAG_NAN2NUM | _cabi.FLAG_WRAP_Y, 0)
    for slot, t in enumerate((ce, cn, ra)):
        lib.plan_set_plane(h, slot, t.data_ptr(), nx, ny * nx, 1)
    lib.plan_set_filter(h, p, 0.1)
    ws = torch.empty(lib.workspace_bytes(h, nbmax), dtype=torch.uint8, device=dev)
    for nb in (int(v) for v in args.nbs.split(",")):
        for lpc in (int(v) for v in args.lpcs.split(",")):
            if lpc > nb:
                continue
            if lpc:
                os.environ["GCMF_FUSED_LEVELS_PER_CTA"] = str(lpc)
            else:
                os.environ.pop("GCMF_FUSED_LEVELS_PER_CTA", None)
            wsb = lib.workspace_bytes(h, nb)
            times = []
            for r in range(args.reps + 1):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                lib.filter(h, nb, [(field.data_ptr(), nx, ny * nx)], [(out.data_ptr(), nx, ny * nx)], ws.data_ptr(), wsb,
                           stream.cuda_stream)
                e1.record(stream)
                torch.cuda.synchronize(dev)
                if r:
                    times.append(e0.elapsed_time(e1))
            ms = min(times)
            print(json.dumps({"nb": nb, "levels_per_cta": lpc, "ms": round(ms, 3),
                              "ms_per_level": round(ms / nb, 4),
                              "gptsteps_per_s": round(nb * ny * nx * args.steps / ms / 1e6, 2)}), flush=True)
    os.environ.pop("GCMF_FUSED_LEVELS_PER_CTA", None)
    lib.plan_destroy(h)


if __name__ == "__main__":
    main()
