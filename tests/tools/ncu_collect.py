"""Copy the ncu summaries of one gpurun call into profiles/ and print one line per kernel (development tool, no GPU).

    python tests/tools/ncu_collect.py ROUND_TAG [gpurun_out]

For every gpurun_out/ncu_<tag>.raw.csv: profiles/ncu_<ROUND_TAG>_<tag>.{raw.csv,details.txt} and a row of
profiles/ncu_<ROUND_TAG>_summary.md (kernel, grid, duration, DRAM bytes read + written, registers, issue-slot and
pipe utilisation, shared-memory wavefronts / bank conflicts).
"""
import csv
import glob
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64_pipe_pct"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
    ("smsp__inst_executed.sum", "warp_insts"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct_of_ncu_peak"),
]

SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0,
         "second": 1.0, "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9}


def read_raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    out = {"kernel": vals[hdr.index("Kernel Name")]}
    for name, key in WANT:
        if name not in hdr:
            continue
        i = hdr.index(name)
        try:
            v = float(vals[i].replace(",", ""))
        except ValueError:
            continue
        out[key] = v * SCALE.get(units[i], 1.0) if key in ("duration", "dram_read", "dram_write") else v
    return out


def main():
    tag = sys.argv[1]
    src = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out")
    lines = ["| capture | kernel | grid x block | regs | duration | DRAM read + write | DRAM GB/s | issue % | fp64 pipe % | "
             "smem wavefronts (bank conflicts) |", "|---|---|---|---|---|---|---|---|---|---|"]
    for raw in sorted(glob.glob(os.path.join(src, "ncu_*.raw.csv"))):
        name = os.path.basename(raw)[len("ncu_"):-len(".raw.csv")]
        try:
            r = read_raw(raw)
        except Exception as exc:  # an empty capture (kernel regex matched nothing)
            print("skipped", raw, exc)
            continue
        for ext in ("raw.csv", "details.txt", "stalls.txt"):
            p = os.path.join(src, f"ncu_{name}.{ext}")
            if os.path.isfile(p):
                shutil.copyfile(p, os.path.join(ROOT, "profiles", f"ncu_{tag}_{name}.{ext}"))
        tot = r.get("dram_read", 0) + r.get("dram_write", 0)
        d = r.get("duration", float("nan"))
        lines.append(
            f"| `ncu_{tag}_{name}` | `{r['kernel'][:90]}` | {int(r.get('grid', 0))} x {int(r.get('block', 0))} | "
            f"{int(r.get('regs', 0))} | {d * 1e3:.3f} ms | {r.get('dram_read', 0) / 1e9:.3f} + {r.get('dram_write', 0) / 1e9:.3f} GB | "
            f"{tot / d / 1e9:.0f} | {r.get('issue_pct', 0):.1f} | {r.get('fp64_pipe_pct', 0):.1f} | "
            f"{r.get('smem_wavefronts', 0) / 1e6:.1f} M ({r.get('smem_bank_conflicts', 0) / 1e6:.1f} M) |")
    text = "\n".join(lines) + "\n"
    with open(os.path.join(ROOT, "profiles", f"ncu_{tag}_summary.md"), "w") as fh:
        fh.write(f"# ncu --set full --clock-control none, one launch per kernel ({tag}; tests/tools/ncu_all.sh)\n\n" + text)
    print(text)


if __name__ == "__main__":
    main()
