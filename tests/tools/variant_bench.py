"""A/B timing of several builds of libgcmf.so in one process (development tool, needs a GPU).

    python tests/tools/variant_bench.py [--nb 62] [--reps 3] name=path/to/lib.so ...

Every library runs the same device-resident cfg3-shaped problem (flux operator, 2400 x 3600 fp64, NaN on the land
quadrant, 44 steps) through gcmf_filter; prints ms per filter call, G grid-point steps/s and the largest
difference of the result from the first library's.
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
    ap.add_argument("--nb", type=int, default=62)
    ap.add_argument("--ny", type=int, default=2400)
    ap.add_argument("--nx", type=int, default=3600)
    ap.add_argument("--steps", type=int, default=44)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--dtype", default="f64")
    ap.add_argument("--op", default="flux", choices=["flux", "reg5"],
                    help="flux: three coefficient planes (cfg3-like); reg5: 5-point stencil with a uint8 wet mask (cfg2-like, "
                         "use --dtype f32 --ny 720 --nx 1440 --nb 365 --steps 11)")
    ap.add_argument("libs", nargs="+")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    tdt = torch.float64 if args.dtype == "f64" else torch.float32
    g = torch.Generator(device=dev).manual_seed(1)
    ny, nx, nb = args.ny, args.nx, args.nb
    field = torch.rand((nb, ny, nx), generator=g, device=dev, dtype=tdt)
    field[:, : ny // 2, : nx // 2] = float("nan")
    wet = torch.ones((ny, nx), device=dev, dtype=tdt)
    wet[: ny // 2, : nx // 2] = 0
    ce = (0.9 + 0.2 * torch.rand((ny, nx), generator=g, device=dev, dtype=tdt)) * wet * torch.roll(wet, -1, 1)
    cn = (0.9 + 0.2 * torch.rand((ny, nx), generator=g, device=dev, dtype=tdt)) * wet * torch.roll(wet, -1, 0)
    ra = 1.0 / (0.9 + 0.2 * torch.rand((ny, nx), generator=g, device=dev, dtype=tdt))
    wet8 = wet.to(torch.uint8).contiguous()
    p = [1.0 / (i + 2) * (-1) ** i for i in range(args.steps + 1)]
    out = torch.empty_like(field)
    first = None
    stream = torch.cuda.current_stream(dev)
    for spec in args.libs:
        name, path = spec.split("=", 1)
        lib = _cabi.Library(os.path.abspath(path))
        gdt = _cabi.GCMF_F64 if tdt == torch.float64 else _cabi.GCMF_F32
        if args.op == "flux":
            h = lib.plan_create(_cabi.OP_FLUX, gdt, ny, nx, _cabi.FLAG_NAN2NUM | _cabi.FLAG_WRAP_Y, 0)
            for slot, t in enumerate((ce, cn, ra)):
                lib.plan_set_plane(h, slot, t.data_ptr(), nx, ny * nx, 1)
        else:
            h = lib.plan_create(_cabi.OP_REGULAR5, gdt, ny, nx,
                                _cabi.FLAG_MASK | _cabi.FLAG_NAN2NUM | _cabi.FLAG_WRAP_Y, 0)
            lib.plan_set_plane(h, 0, wet8.data_ptr(), nx, ny * nx, 1)
        lib.plan_set_filter(h, p, 0.1)
        wsb = lib.workspace_bytes(h, nb)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        fin = [(field.data_ptr(), nx, ny * nx)]
        fout = [(out.data_ptr(), nx, ny * nx)]
        times = []
        for r in range(args.reps + 1):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            lib.filter(h, nb, fin, fout, ws.data_ptr(), wsb, stream.cuda_stream)
            e1.record(stream)
            torch.cuda.synchronize(dev)
            if r:
                times.append(e0.elapsed_time(e1))
        ms = min(times)
        res = torch.nan_to_num(out, nan=0.0)
        if first is None:
            first = res.clone()
            diff = 0.0
        else:
            diff = float((res - first).abs().max())
        print(json.dumps({"variant": name, "ms": round(ms, 3), "ms_all": [round(t, 3) for t in times],
                          "gptsteps_per_s": round(nb * ny * nx * args.steps / ms / 1e6, 2), "maxdiff_vs_first": diff,
                          "nan_out": int(torch.isnan(out).sum())}), flush=True)
        lib.plan_destroy(h)
        del ws


if __name__ == "__main__":
    main()
