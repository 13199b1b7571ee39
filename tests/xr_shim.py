"""A tiny stand-in for the parts of xarray that gcm-filters touches (xarray is not installed in this
image): named-dimension DataArray / Variable / Dataset and apply_ufunc with core dims.  TEST
INFRASTRUCTURE ONLY -- it lets the Filter.apply / apply_to_vector xarray branch be exercised; where real
xarray exists the same code path runs against it."""
import copy
import sys
import types

import numpy as np


class Variable:
    def __init__(self, data, dims=None, name=None):
        self.data = np.asarray(data)
        self.dims = tuple(dims) if dims is not None else tuple(f"dim_{k}" for k in range(self.data.ndim))
        assert len(self.dims) == self.data.ndim
        self.name = name

    @property
    def values(self):
        return self.data

    @property
    def dtype(self):
        return self.data.dtype

    @property
    def shape(self):
        return self.data.shape

    def mean(self, dim):
        dims = (dim,) if isinstance(dim, str) else tuple(dim)
        axes = tuple(self.dims.index(d) for d in dims)
        return DataArray(np.nanmean(self.data, axis=axes), [d for d in self.dims if d not in dims])


class DataArray(Variable):
    pass


class Dataset:
    def __init__(self, data_vars=None):
        self.variables = {}
        for k, v in (data_vars or {}).items():
            self[k] = v

    def __setitem__(self, key, value):
        if not isinstance(value, Variable):
            value = DataArray(value)
        self.variables[key] = value

    def __getitem__(self, key):
        return self.variables[key]

    def copy(self, deep=True):
        return copy.deepcopy(self) if deep else copy.copy(self)


def apply_ufunc(func, *args, input_core_dims, output_core_dims, output_dtypes=None, dask=None):
    """Core dims are moved to the end (in the order given), the remaining dims are aligned by name and
    broadcast through size-1 axes, exactly what the reference relies on (filter.py:478-486)."""
    batch = []
    for a, core in zip(args, input_core_dims):
        for d in a.dims:
            if d not in core and d not in batch:
                batch.append(d)
    arrays = []
    for a, core in zip(args, input_core_dims):
        own_batch = [d for d in batch if d in a.dims]
        order = [a.dims.index(d) for d in own_batch + list(core)]
        arr = np.transpose(a.data, order)
        shape = [a.data.shape[a.dims.index(d)] if d in a.dims else 1 for d in batch]
        shape += [a.data.shape[a.dims.index(d)] for d in core]
        arrays.append(arr.reshape(shape))
    res = func(*arrays)
    single = not isinstance(res, tuple)
    res = (res,) if single else res
    outs = tuple(DataArray(r, batch + list(core)) for r, core in zip(res, output_core_dims))
    return outs[0] if single else outs


_previous = None


def install():
    """Register the shim as the module `xarray` unless the real package is importable.  (The oracle's
    reference loader may have left its own minimal stub there: it is put back by uninstall().)"""
    global _previous
    existing = sys.modules.get("xarray")
    if existing is None:
        try:
            import xarray as xr
            return xr
        except ImportError:
            pass
    elif hasattr(existing, "apply_ufunc") and not getattr(existing, "__gcmf_shim__", False):
        return existing  # the real xarray
    _previous = existing if existing is not None and not getattr(existing, "__gcmf_shim__", False) else _previous
    mod = types.ModuleType("xarray")
    mod.Variable, mod.DataArray, mod.Dataset, mod.apply_ufunc = Variable, DataArray, Dataset, apply_ufunc
    mod.__gcmf_shim__ = True
    sys.modules["xarray"] = mod
    return mod


def uninstall():
    global _previous
    mod = sys.modules.get("xarray")
    if mod is not None and getattr(mod, "__gcmf_shim__", False):
        if _previous is not None:
            sys.modules["xarray"] = _previous
        else:
            del sys.modules["xarray"]
    _previous = None
