"""Build libgcmf.so in-tree with nvcc for sm_100a.  `python -m gcm_filters_b200.build [--force]`."""
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
OUT = os.path.join(PKG, "libgcmf.so")
SOURCES = ["gcmf.cu"]
HEADERS = ["gcmf_stencils.cuh", "gcmf_fused.cuh", "gcmf_march.cuh", "gcmf_vec2.cuh", "gcmf_internal.h",
           os.path.join("..", "..", "include", "gcmf.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",  # separate rounding of * and +, in the reference's evaluation order
    "-shared", "-Xcompiler", "-fPIC", "-Xptxas", "-v", "-diag-suppress", "128",
]


HASH = OUT + ".srchash"


def source_hash(files=None):
    """sha256 over the CUDA sources, headers and flags: identifies what a built libgcmf.so was made from.
    `files`: hash only these files of csrc/ (what one kernel family is made from; profiles/traffic.json uses it to tell
    whether a committed ncu capture still describes the kernel in the tree)."""
    import hashlib

    h = hashlib.sha256()
    for rel in (SOURCES + HEADERS if files is None else list(files)):
        with open(os.path.join(CSRC, rel), "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_stale():
    """True when libgcmf.so exists but was built from different sources than the ones in the tree."""
    if not os.path.isfile(OUT) or not os.path.isfile(HASH):
        return False
    try:
        with open(HASH) as fh:
            return fh.read().strip() != source_hash()
    except OSError:
        return False


def find_nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found")


def up_to_date():
    if not os.path.isfile(OUT) or not os.path.isfile(HASH):
        return False
    with open(HASH) as fh:
        return fh.read().strip() == source_hash()


def build(force=False, verbose=False):
    if not force and up_to_date():
        return OUT
    cmd = [find_nvcc()] + NVCC_FLAGS + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", OUT]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    with open(HASH, "w") as fh:
        fh.write(source_hash())
    if verbose:
        print(res.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
