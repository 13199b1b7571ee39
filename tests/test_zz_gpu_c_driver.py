"""GPU test, sorted last on purpose: the plain-C driver of the C ABI (tests/cabi/gpu_vs_emu.c) runs every operator
family through libgcmf.so on the GPU and through the host emulator of the same sources and compares bit for bit.
No Python on the compute path: this is the drop-in boundary exactly as a C / Fortran / Julia caller would use it."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# measured bit-identical on a B200 at the end of round 1 (profiles/gpu_vs_emu_r01.log)
VERIFIED = ["cgrid f64 37x54", "cgrid f32 20x24", "cgrid f64 7x70", "reg5 masked f64 40x264 fused",
            "reg5 masked f32 40x264 fused", "reg5 unmasked f64 36x128 fused", "reg5 masked f64 band",
            "reg5 masked f32 band", "flux f64 48x256 fused (control)"]


@pytest.mark.gpu
def test_c_abi_on_the_gpu_is_bit_identical_to_the_emulator(tmp_path):
    cuda_inc = "/usr/local/cuda/include"
    if shutil.which("gcc") is None or not os.path.isfile(os.path.join(cuda_inc, "cuda_runtime_api.h")):
        pytest.skip("needs gcc and the CUDA runtime headers")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from gcm_filters_b200 import _cabi
    from hostemu_util import emu_library
    _cabi.get_library()  # builds libgcmf.so if it is missing
    emu_library()        # builds the emulator if it is missing
    exe = str(tmp_path / "gpu_vs_emu")
    subprocess.run(["gcc", "-std=c99", "-O1", os.path.join(ROOT, "tests", "cabi", "gpu_vs_emu.c"),
                    "-I", os.path.join(ROOT, "include"), "-I", cuda_inc, "-L", "/usr/local/cuda/lib64",
                    "-Wl,-rpath,/usr/local/cuda/lib64", "-lcudart", "-ldl", "-lm", "-o", exe], check=True)
    res = subprocess.run([exe], cwd=ROOT, capture_output=True, text=True, timeout=300)
    print(res.stdout)
    lines = [l for l in res.stdout.splitlines() if "values" in l or "FAILED" in l]
    assert "sm_arch 100" in res.stdout and len(lines) >= len(VERIFIED), res.stdout + res.stderr
    assert "CUDA error" not in res.stdout and "FAILED TO RUN" not in res.stdout, res.stdout
    for name in VERIFIED:
        hit = [l for l in lines if l.startswith(name)]
        assert hit and hit[0].rstrip().endswith("bit-identical"), hit or name
    # the remaining cases (added after the round's GPU budget was spent) must agree to rounding at least; a line
    # that is not bit-identical is reported in the captured output above
    for l in lines:
        if not l.rstrip().endswith("bit-identical"):
            worst = float(l.split("worst relative difference")[1].split()[0])
            nan_mismatch = "differ" in l and worst == 0.0
            assert worst < 1e-12 and not nan_mismatch, l
