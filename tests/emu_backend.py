"""TEST-ONLY: run the public API (Filter.apply, Laplacian.__call__, prepare / finalize) against the host
emulator of libgcmf (tests/hostemu: the same CUDA source compiled with g++, every launch a host loop).

``install(monkeypatch)`` swaps the three entry points of ``gcm_filters_b200.engine`` that touch the device for
numpy versions that drive the emulator through the same C ABI, so that one test body can run on the CPU (host
logic, plane precombination, index handling, step sequencing) and, marked ``gpu``, on the B200 (the product).
The product never imports this module and has no CPU path of its own.
"""
import numpy as np

from gcm_filters_b200 import engine

from hostemu_util import EmuPlan


def _plan(lap, np_dtype, ny, nx):
    key = ("hostemu", np.dtype(np_dtype).str, ny, nx)
    plan = lap._device_state.get(key)
    if plan is None:
        plan = lap._device_state[key] = EmuPlan(lap, np_dtype, ny, nx)
    return plan


def _stage(lap, fields):
    if len(fields) != lap.ncomp:
        raise ValueError(f"expected {lap.ncomp} field component(s), got {len(fields)}")
    arrs = [np.asarray(getattr(f, "values", f)) for f in fields]
    shape = arrs[0].shape
    if len(shape) < 2:
        raise ValueError("fields need at least two dimensions (y, x)")
    if any(a.shape != shape for a in arrs[1:]):
        raise ValueError("vector components must have the same shape")
    dt = lap.compute_dtype(engine._float_dtype_of(arrs[0]))
    return arrs, _plan(lap, dt, shape[-2], shape[-1])


def run_laplacian(lap, fields):
    arrs, plan = _stage(lap, fields)
    return tuple(plan.laplacian(arrs))


def run_filter(lap, p, c, fields, out=None):
    arrs, plan = _stage(lap, fields)
    res = plan.filter(arrs, [float(v) for v in p], float(c))
    if out is not None:
        out = tuple(out) if isinstance(out, (tuple, list)) else (out,)
        for dst, r in zip(out, res):
            dst[...] = r
        return out
    return tuple(res)


def run_area_op(lap, field, divide):
    arrs, plan = _stage(lap, (field,))
    fin = [np.ascontiguousarray(arrs[0], dtype=plan.dtype).reshape((-1, plan.ny, plan.nx))]
    out = [np.full_like(fin[0], 777.0)]
    op = plan.lib.finalize if divide else plan.lib.prepare
    op(plan.h, fin[0].shape[0], plan._specs(fin), plan._specs(out))
    return out[0].reshape(arrs[0].shape)


def install(monkeypatch):
    monkeypatch.setattr(engine, "run_laplacian", run_laplacian)
    monkeypatch.setattr(engine, "run_filter", run_filter)
    monkeypatch.setattr(engine, "run_area_op", run_area_op)
