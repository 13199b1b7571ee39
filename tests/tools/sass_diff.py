"""Development tool: which device functions of the in-tree libgcmf.so differ from a build of commit REF?

    python tests/tools/sass_diff.py REF          (e.g. the last commit whose build passed `pytest -m gpu`)

Used when code is changed without a GPU at hand (opt-in GCMF_OPT_* variants that default to off, host-side fixes):
functions whose SASS is unchanged are still covered by the last GPU verification.
"""
import hashlib
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from gcm_filters_b200 import build as b  # noqa: E402


def functions(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    table, cur, buf = {}, None, []
    for line in out.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            if cur:
                table[cur] = hashlib.md5("\n".join(buf).encode()).hexdigest()
            cur, buf = m.group(1), []
        elif cur:
            buf.append(line)
    if cur:
        table[cur] = hashlib.md5("\n".join(buf).encode()).hexdigest()
    return table


def main():
    ref = sys.argv[1]
    tmp = tempfile.mkdtemp()
    tar = subprocess.run(["git", "-C", ROOT, "archive", ref, "gcm_filters_b200/csrc", "include"], capture_output=True, check=True)
    subprocess.run(["tar", "x", "-C", tmp], input=tar.stdout, check=True)
    flags = [f for f in b.NVCC_FLAGS if f not in ("-Xptxas", "-v")]
    subprocess.run([b.find_nvcc()] + flags + ["gcm_filters_b200/csrc/gcmf.cu", "-o", "ref.so"], cwd=tmp, check=True,
                   capture_output=True)
    b.build()
    old, new = functions(os.path.join(tmp, "ref.so")), functions(b.OUT)
    changed = sorted(k for k in set(old) | set(new) if old.get(k) != new.get(k))
    names = subprocess.run(["c++filt"], input="\n".join(changed), capture_output=True, text=True).stdout.split("\n")
    print(f"{len(changed)} of {len(new)} device functions differ from {ref}")
    for n in sorted(set(re.sub(r"\(.*", "", n) for n in names if n)):
        print("  ", n)


if __name__ == "__main__":
    main()
