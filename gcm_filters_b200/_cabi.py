"""ctypes binding of libgcmf.so (C ABI declared in include/gcmf.h).

The product always goes through :func:`get_library`, which loads the in-tree CUDA library and
fails loudly when it is missing -- there is no CPU fallback.  (Tests may construct
:class:`Library` on another path to drive the test-only host emulator of the same sources.)
"""
import ctypes
import os
import threading

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG_DIR, "libgcmf.so")

GCMF_F32, GCMF_F64 = 0, 1
OP_REGULAR5, OP_FLUX, OP_VECTOR_B, OP_VECTOR_C = 0, 1, 2, 3
FLAG_MASK, FLAG_NAN2NUM, FLAG_FOLD_N, FLAG_CUT_S, FLAG_WRAP_Y, FLAG_AREA = 1, 2, 4, 8, 16, 32

EXPORTS = [
    "gcmf_version", "gcmf_sm_arch", "gcmf_last_error", "gcmf_plan_create", "gcmf_plan_destroy",
    "gcmf_plan_set_plane", "gcmf_plan_set_filter", "gcmf_workspace_bytes", "gcmf_laplacian", "gcmf_filter",
    "gcmf_cheb_step", "gcmf_prepare", "gcmf_launch_count", "gcmf_fused_max_steps", "gcmf_plan_set_steps_per_block",
    "gcmf_cheb_fused", "gcmf_cheb_step_halo", "gcmf_halo_push", "gcmf_finalize", "gcmf_cheb_fused_halo",
]


class PlanDesc(ctypes.Structure):
    _fields_ = [("op", ctypes.c_int32), ("dtype", ctypes.c_int32), ("ny", ctypes.c_int32), ("nx", ctypes.c_int32),
                ("flags", ctypes.c_int32), ("device", ctypes.c_int32)]


class Field(ctypes.Structure):
    _fields_ = [("ptr", ctypes.c_void_p), ("pitch", ctypes.c_int64), ("bstride", ctypes.c_int64)]


class Halo(ctypes.Structure):
    _fields_ = [("north_ghost", ctypes.c_void_p * 2), ("south_ghost", ctypes.c_void_p * 2),
                ("north_bstride", ctypes.c_int64), ("south_bstride", ctypes.c_int64),
                ("wait_north", ctypes.c_void_p), ("wait_south", ctypes.c_void_p),
                ("signal_north", ctypes.c_void_p), ("signal_south", ctypes.c_void_p),
                ("wait_value", ctypes.c_uint32), ("signal_value", ctypes.c_uint32), ("counters", ctypes.c_void_p)]


class GcmfError(RuntimeError):
    pass


class Library:
    """Typed handle on one build of the gcmf C ABI."""

    def __init__(self, path):
        if not os.path.isfile(path):
            raise GcmfError(
                f"{path} not found: the CUDA extension has not been built. Run "
                "`python -c 'import __graft_entry__ as g; g.build()'` (or `python -m gcm_filters_b200.build`). "
                "gcm_filters_b200 has no CPU fallback.")
        self.path = path
        self.lib = lib = ctypes.CDLL(path)
        vp, i32, i64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64
        fp = ctypes.POINTER(Field)
        lib.gcmf_version.restype = ctypes.c_int
        lib.gcmf_sm_arch.restype = ctypes.c_int
        lib.gcmf_last_error.restype = ctypes.c_char_p
        lib.gcmf_launch_count.restype = i64
        lib.gcmf_plan_create.argtypes = [ctypes.POINTER(PlanDesc), ctypes.POINTER(vp)]
        lib.gcmf_plan_destroy.argtypes = [vp]
        lib.gcmf_plan_set_plane.argtypes = [vp, ctypes.c_int, vp, i64, i64, i32]
        lib.gcmf_plan_set_filter.argtypes = [vp, i32, ctypes.POINTER(ctypes.c_double), ctypes.c_double]
        lib.gcmf_workspace_bytes.argtypes = [vp, i64, ctypes.POINTER(ctypes.c_size_t)]
        lib.gcmf_laplacian.argtypes = [vp, i64, fp, fp, vp]
        lib.gcmf_prepare.argtypes = [vp, i64, fp, fp, vp]
        lib.gcmf_finalize.argtypes = [vp, i64, fp, fp, vp]
        lib.gcmf_finalize.restype = ctypes.c_int
        lib.gcmf_filter.argtypes = [vp, i64, fp, fp, vp, ctypes.c_size_t, vp]
        lib.gcmf_cheb_step.argtypes = [vp, i64, i32, fp, fp, fp, fp, vp]
        lib.gcmf_fused_max_steps.argtypes = [vp]
        lib.gcmf_fused_max_steps.restype = ctypes.c_int
        lib.gcmf_plan_set_steps_per_block.argtypes = [vp, i32]
        lib.gcmf_plan_set_steps_per_block.restype = ctypes.c_int
        lib.gcmf_cheb_fused.argtypes = [vp, i64, i32, i32, fp, fp, fp, fp, fp, vp]
        lib.gcmf_cheb_fused.restype = ctypes.c_int
        hp = ctypes.POINTER(Halo)
        lib.gcmf_cheb_step_halo.argtypes = [vp, i64, i32, fp, fp, fp, fp, hp, vp]
        lib.gcmf_cheb_step_halo.restype = ctypes.c_int
        lib.gcmf_halo_push.argtypes = [vp, i64, fp, hp, vp]
        lib.gcmf_halo_push.restype = ctypes.c_int
        lib.gcmf_cheb_fused_halo.argtypes = [vp, i64, i32, i32, fp, fp, fp, fp, fp, hp, hp, vp]
        lib.gcmf_cheb_fused_halo.restype = ctypes.c_int
        for name in ("gcmf_plan_create", "gcmf_plan_destroy", "gcmf_plan_set_plane", "gcmf_plan_set_filter",
                     "gcmf_workspace_bytes", "gcmf_laplacian", "gcmf_prepare", "gcmf_filter", "gcmf_cheb_step"):
            getattr(lib, name).restype = ctypes.c_int

    def check(self, rc):
        if rc != 0:
            raise GcmfError(f"libgcmf error {rc}: {self.lib.gcmf_last_error().decode()}")

    # -- thin wrappers -------------------------------------------------------------------
    def plan_create(self, op, dtype, ny, nx, flags, device):
        desc = PlanDesc(op, dtype, ny, nx, flags, device)
        h = ctypes.c_void_p()
        self.check(self.lib.gcmf_plan_create(ctypes.byref(desc), ctypes.byref(h)))
        return h

    def plan_destroy(self, h):
        self.lib.gcmf_plan_destroy(h)

    def plan_set_plane(self, h, slot, ptr, pitch, bstride, nb):
        self.check(self.lib.gcmf_plan_set_plane(h, slot, ctypes.c_void_p(ptr), pitch, bstride, nb))

    def plan_set_filter(self, h, p, c):
        n = len(p) - 1
        arr = (ctypes.c_double * len(p))(*[float(v) for v in p])
        self.check(self.lib.gcmf_plan_set_filter(h, n, arr, float(c)))

    def workspace_bytes(self, h, nb):
        out = ctypes.c_size_t()
        self.check(self.lib.gcmf_workspace_bytes(h, nb, ctypes.byref(out)))
        return out.value

    @staticmethod
    def fields(specs):
        """specs: list of (ptr, pitch, bstride) -> ctypes array of gcmf_field (or None)."""
        if specs is None or isinstance(specs, ctypes.Array):  # already converted (callers on a launch-bound path cache them)
            return specs
        arr = (Field * len(specs))()
        for k, (ptr, pitch, bstride) in enumerate(specs):
            arr[k] = Field(ptr, pitch, bstride)
        return arr

    def laplacian(self, h, nb, fin, fout, stream=0):
        self.check(self.lib.gcmf_laplacian(h, nb, self.fields(fin), self.fields(fout), ctypes.c_void_p(stream)))

    def prepare(self, h, nb, fin, fout, stream=0):
        self.check(self.lib.gcmf_prepare(h, nb, self.fields(fin), self.fields(fout), ctypes.c_void_p(stream)))

    def finalize(self, h, nb, fin, fout, stream=0):
        self.check(self.lib.gcmf_finalize(h, nb, self.fields(fin), self.fields(fout), ctypes.c_void_p(stream)))

    def filter(self, h, nb, fin, fout, ws_ptr, ws_bytes, stream=0):
        self.check(self.lib.gcmf_filter(h, nb, self.fields(fin), self.fields(fout), ctypes.c_void_p(ws_ptr),
                                        ws_bytes, ctypes.c_void_p(stream)))

    def cheb_step(self, h, nb, step, t1, t2, t0, bar, stream=0):
        self.check(self.lib.gcmf_cheb_step(h, nb, step, self.fields(t1), self.fields(t2), self.fields(t0),
                                           self.fields(bar), ctypes.c_void_p(stream)))

    def fused_max_steps(self, h):
        return int(self.lib.gcmf_fused_max_steps(h))

    def set_steps_per_block(self, h, k):
        self.check(self.lib.gcmf_plan_set_steps_per_block(h, k))

    def cheb_fused(self, h, nb, step, k, t1, t2, t1o, t2o, bar, stream=0):
        self.check(self.lib.gcmf_cheb_fused(h, nb, step, k, self.fields(t1), self.fields(t2), self.fields(t1o),
                                            self.fields(t2o), self.fields(bar), ctypes.c_void_p(stream)))

    def cheb_step_halo(self, h, nb, step, t1, t2, t0, bar, halo, stream=0):
        self.check(self.lib.gcmf_cheb_step_halo(h, nb, step, self.fields(t1), self.fields(t2), self.fields(t0),
                                                self.fields(bar), ctypes.byref(halo), ctypes.c_void_p(stream)))

    def cheb_fused_halo(self, h, nb, step, k, t1, t2, t1o, t2o, bar, halo1, halo2, stream=0):
        self.check(self.lib.gcmf_cheb_fused_halo(h, nb, step, k, self.fields(t1), self.fields(t2), self.fields(t1o),
                                                 self.fields(t2o), self.fields(bar), ctypes.byref(halo1),
                                                 ctypes.byref(halo2), ctypes.c_void_p(stream)))

    def halo_push(self, h, nb, field, halo, stream=0):
        self.check(self.lib.gcmf_halo_push(h, nb, self.fields(field), ctypes.byref(halo), ctypes.c_void_p(stream)))

    def launch_count(self):
        return int(self.lib.gcmf_launch_count())


_lib = None
_lock = threading.Lock()


def get_library():
    """The in-tree CUDA build of libgcmf.so; raises GcmfError if it has not been built."""
    global _lib
    with _lock:
        if _lib is None:
            from . import build as _build

            if not os.path.isfile(LIB_PATH) or _build.is_stale():
                # missing, or built from other sources than the ones in the tree: rebuild in place with nvcc
                # (about 45 s); without a toolchain this raises -- there is no fallback to anything else
                try:
                    _build.build(force=True)
                except Exception as exc:
                    raise GcmfError(f"{LIB_PATH} is missing or out of date and could not be rebuilt ({exc}); "
                                    "run `python -m gcm_filters_b200.build`. gcm_filters_b200 has no CPU fallback.")
            lib = Library(LIB_PATH)
            if lib.lib.gcmf_sm_arch() != 100:
                raise GcmfError(f"{LIB_PATH} is not an sm_100a build")
            _lib = lib
        return _lib
