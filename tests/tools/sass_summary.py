"""Per-kernel SASS opcode summary of libgcmf.so (development tool; cuobjdump only, no GPU).

    python tests/tools/sass_summary.py [lib.so] > profiles/sass_rNN_summary.txt

For every kernel: instruction count, the TMA / mbarrier / tensor-core evidence mnemonics (UTMALDG = cp.async.bulk.tensor,
UBLKCP = cp.async.bulk, SYNCS = mbarrier, LDGSTS = cp.async; UTC*MMA / LDTM / HMMA would be tensor cores: none expected,
the path is a memory-bound stencil), local-memory traffic (LDL / STL = spills) and the top opcodes.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gcm_filters_b200", "libgcmf.so")
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    demangle = {}
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.search(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur:
            kernels[cur][m.group(1)] += 1
    try:
        names = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
        demangle = dict(zip(kernels, names))
    except OSError:
        pass
    watch = ["UTMALDG", "UBLKCP", "SYNCS", "LDGSTS", "UTCHMMA", "HMMA", "LDTM", "LDL", "STL", "LDS", "STS", "LDG", "STG",
             "DFMA", "DADD", "DMUL", "FFMA", "SHFL", "BAR"]
    print(f"# SASS summary of {os.path.relpath(lib, ROOT)} (sm_100a), {len(kernels)} kernels\n")
    for k, c in kernels.items():
        base = collections.Counter()
        for op, n in c.items():
            base[op.split(".")[0]] += n
        total = sum(c.values())
        print(f"{demangle.get(k, k)}")
        print(f"    instructions {total}; " + ", ".join(f"{w} {base[w]}" for w in watch if base[w]))
        print("    top: " + ", ".join(f"{op} {n}" for op, n in base.most_common(8)) + "\n")


if __name__ == "__main__":
    main()
