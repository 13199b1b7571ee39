"""TEST-ONLY driver of the host emulator (tests/hostemu): the same C ABI and the same stencil source
as libgcmf.so, compiled with g++ so that every kernel launch is a host loop.  Lets the CPU test
suite check index handling, plane precombination and step sequencing without a GPU."""
import os
import subprocess

import numpy as np

from gcm_filters_b200 import _cabi

HERE = os.path.dirname(os.path.abspath(__file__))
EMU_DIR = os.path.join(HERE, "hostemu")
EMU_LIB = os.path.join(EMU_DIR, "libgcmf_hostemu.so")
_DT = {np.dtype(np.float32): _cabi.GCMF_F32, np.dtype(np.float64): _cabi.GCMF_F64}
_lib = None


def emu_library():
    global _lib
    if _lib is None and os.environ.get("GCMF_HOSTEMU_LIB"):  # e.g. an ASAN / UBSAN build (tests/tools/asan_fuzz.sh)
        _lib = _cabi.Library(os.environ["GCMF_HOSTEMU_LIB"])
        assert _lib.lib.gcmf_sm_arch() == 0
    if _lib is None:
        src = os.path.join(HERE, "..", "gcm_filters_b200", "csrc")
        deps = [os.path.join(src, f) for f in ("gcmf.cu", "gcmf_stencils.cuh", "gcmf_internal.h")]
        if not os.path.isfile(EMU_LIB) or any(os.path.getmtime(d) > os.path.getmtime(EMU_LIB) for d in deps):
            subprocess.run(["sh", os.path.join(EMU_DIR, "build.sh")], check=True)
        _lib = _cabi.Library(EMU_LIB)
        assert _lib.lib.gcmf_sm_arch() == 0
    return _lib


class EmuPlan:
    def __init__(self, lap, dtype, ny, nx, flags_override=None):
        self.lib = emu_library()
        self.dtype = np.dtype(dtype)
        spec = lap._planes
        flags = spec.flags if flags_override is None else flags_override
        self.h = self.lib.plan_create(spec.op, _DT[self.dtype], ny, nx, flags, 0)
        self.keep = []
        for slot, pl in enumerate(spec.planes):
            if slot == 0 and spec.op == _cabi.OP_REGULAR5:
                if spec.mask is None:
                    continue
                a = np.ascontiguousarray(spec.mask, dtype=np.uint8)
            elif pl is None:
                continue
            else:
                a = np.ascontiguousarray(pl, dtype=self.dtype)
            a = a.reshape((-1, ny, nx))
            self.keep.append(a)
            self.lib.plan_set_plane(self.h, slot, a.ctypes.data, nx, ny * nx, a.shape[0])
        self.ncomp = lap.ncomp
        self.ny, self.nx = ny, nx

    def _specs(self, arrs):
        return [(a.ctypes.data, self.nx, self.ny * self.nx) for a in arrs]

    def laplacian(self, fields):
        fin = [np.ascontiguousarray(f, dtype=self.dtype).reshape((-1, self.ny, self.nx)) for f in fields]
        out = [np.full_like(f, 777.0) for f in fin]
        self.lib.laplacian(self.h, fin[0].shape[0], self._specs(fin), self._specs(out))
        return [o.reshape(np.shape(fields[0])) for o in out]

    def filter(self, fields, p, c):
        fin = [np.ascontiguousarray(f, dtype=self.dtype).reshape((-1, self.ny, self.nx)) for f in fields]
        keep = [f.copy() for f in fin]
        out = [np.full_like(f, 777.0) for f in fin]
        nb = fin[0].shape[0]
        self.lib.plan_set_filter(self.h, p, c)
        nbytes = self.lib.workspace_bytes(self.h, nb)
        raw = np.zeros(nbytes + 256, dtype=np.uint8)
        off = (-raw.ctypes.data) % 256
        self.lib.filter(self.h, nb, self._specs(fin), self._specs(out), raw.ctypes.data + off, nbytes)
        for a, b in zip(fin, keep):
            assert np.array_equal(a, b, equal_nan=True), "input was modified"
        return [o.reshape(np.shape(fields[0])) for o in out]


def emu_set_steps_per_block(plan, k):
    plan.lib.set_steps_per_block(plan.h, k)
