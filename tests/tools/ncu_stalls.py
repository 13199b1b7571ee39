"""Source-level stall summary of one `ncu --set full --import-source on` capture (development tool, no GPU).

    python tests/tools/ncu_stalls.py gpurun_out/ncu_TAG.source.csv "title" > profiles/ncu_rNN_TAG.stalls.txt

Input: `ncu -i X.ncu-rep --page source --csv`.  Output: warp-stall samples by reason, executed warp instructions by
opcode, and the instructions with the most samples (with their top stall reasons).
"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else path
    rows = list(csv.reader(open(path)))
    kernel = rows[0][1] if rows and rows[0] and rows[0][0] == "Kernel Name" else ""
    hdr = rows[1] if kernel else rows[0]
    body = rows[2:] if kernel else rows[1:]
    ci, cs, ce = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "(Not Issued)" not in h]

    def num(v):
        try:
            return float(v)
        except ValueError:
            return 0.0

    reasons = collections.Counter()
    ops, op_samples = collections.Counter(), collections.Counter()
    insts = []
    for r in body:
        if len(r) <= max(ci, cs, ce):
            continue
        n, ex = num(r[cs]), num(r[ce])
        m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", r[ci])
        op = m.group(1) if m else "?"
        ops[op] += ex
        op_samples[op] += n
        per = [(h, num(r[i])) for i, h in stall_cols]
        for h, v in per:
            reasons[h] += v
        insts.append((n, ex, r[ci].strip(), sorted(per, key=lambda t: -t[1])[:3]))
    tot_s, tot_e = sum(reasons.values()), sum(ops.values())
    print(title)
    print(kernel)
    print(f"total warp-stall samples {tot_s:.0f}; executed warp instructions {tot_e:.0f}\n")
    print("stall reason totals:")
    for h, v in reasons.most_common():
        if v:
            print(f"  {h:28s} {v:8.0f} {100 * v / tot_s:5.1f}%")
    print("\nexecuted warp instructions by opcode (share of executed, share of samples):")
    for op, v in ops.most_common(24):
        print(f"  {op:10s} {v:12.0f} {100 * v / tot_e:5.1f}%   samples {100 * op_samples[op] / max(tot_s, 1):5.1f}%")
    print("\ntop instructions by samples (samples, share, executed, SASS, top stall reasons):")
    for n, ex, src, per in sorted(insts, key=lambda t: -t[0])[:40]:
        print(f" {n:6.0f} {100 * n / max(tot_s, 1):4.1f}% ex={ex:10.0f} {src[:70]:70s} {[(h, int(v)) for h, v in per if v]}")


if __name__ == "__main__":
    main()
