"""Build an A/B variant of libgcmf.so with extra -D switches (development tool; nvcc cross-compiles without a GPU).

    python tests/tools/build_variant.py NAME [-DGCMF_OPT_BARPF=1 ...]   ->  build/variants/libgcmf_NAME.so

The variants travel to the GPU box with the snapshot (`*.so` is git-ignored, not gpurun-ignored) and are timed
against each other in one process by tests/tools/variant_bench.py.  The product only ever loads
gcm_filters_b200/libgcmf.so.
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from gcm_filters_b200 import build as b  # noqa: E402


def main():
    name, defs = sys.argv[1], sys.argv[2:]
    out_dir = os.path.join(ROOT, "build", "variants")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, f"libgcmf_{name}.so")
    cmd = [b.find_nvcc()] + b.NVCC_FLAGS + defs + [os.path.join(b.CSRC, s) for s in b.SOURCES] + ["-o", out]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.exit("nvcc failed:\n" + res.stdout + res.stderr)
    with open(out + ".ptxas.log", "w") as fh:
        fh.write(res.stderr)
    print(out)


if __name__ == "__main__":
    main()
