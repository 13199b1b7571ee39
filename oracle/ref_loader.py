"""Import the live reference (``/root/reference/gcm_filters``) on plain numpy.

TEST INFRASTRUCTURE ONLY.  The reference imports xarray at module scope
(``gcm_filters/filter.py:10``) and builds an ``xr.Dataset`` in ``Filter.__post_init__``
(``filter.py:393``); xarray is not installed here, so a tiny stub module is injected.
``_create_filter_func(spec, Laplacian)(field, *planes)`` is exactly the callable that
``xr.apply_ufunc`` would invoke (``filter.py:474-480``), so calling it directly exercises
the reference's hot path byte for byte.

The reference tree exists only in the build container: everything here raises
``ReferenceUnavailable`` elsewhere and callers (tests) skip.
"""
import importlib
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
# where an UNMODIFIED copy of the reference package may live: an explicit override, the build container's read-only
# tree, or the `pip install --target baseline/_ref /root/reference` copy that travels to the GPU box (git-ignored)
CANDIDATES = [os.environ.get("GCMF_REFERENCE"), "/root/reference", os.path.join(os.path.dirname(_HERE), "baseline", "_ref")]


class ReferenceUnavailable(RuntimeError):
    pass


def reference_root():
    for root in CANDIDATES:
        if root and os.path.isfile(os.path.join(root, "gcm_filters", "kernels.py")):
            return root
    return None


REFERENCE_ROOT = reference_root() or "/root/reference"


def available():
    return reference_root() is not None


def _xarray_stub():
    mod = types.ModuleType("xarray")

    class Dataset(dict):
        """dict of name -> array, enough for ``filter.py:393`` and ``grid_ds[name]``."""

    class DataArray:  # placeholder: only referenced in isinstance checks
        pass

    mod.Dataset = Dataset
    mod.DataArray = DataArray
    mod.__gcmf_stub__ = True
    return mod


_ref = None


def load():
    """Return the imported reference package ``gcm_filters`` (cached)."""
    global _ref
    if _ref is not None:
        return _ref
    root = reference_root()
    if root is None:
        raise ReferenceUnavailable(f"no reference tree in {[c for c in CANDIDATES if c]}")
    sys.dont_write_bytecode = True  # the tree is read-only
    if "xarray" not in sys.modules:
        try:
            importlib.import_module("xarray")
        except ImportError:
            sys.modules["xarray"] = _xarray_stub()
    if root not in sys.path:
        sys.path.insert(0, root)
    _ref = importlib.import_module("gcm_filters")
    return _ref


def ref_laplacian(grid_type_name, grid_vars, *fields):
    """One application of the reference Laplacian (``kernels.py`` ``__call__``)."""
    gf = load()
    kernels = importlib.import_module("gcm_filters.kernels")
    gt = kernels.GridType[grid_type_name]
    lap = kernels.ALL_KERNELS[gt](**{k: v.copy() for k, v in grid_vars.items()})
    return lap(*fields)


def ref_filter(grid_type_name, grid_vars, fields, **filter_args):
    """Full reference filter through ``Filter`` + ``_create_filter_func[_vec]``."""
    gf = load()
    kernels = importlib.import_module("gcm_filters.kernels")
    fmod = importlib.import_module("gcm_filters.filter")
    gt = kernels.GridType[grid_type_name]
    if "filter_shape" in filter_args and isinstance(filter_args["filter_shape"], str):
        filter_args = dict(filter_args)
        filter_args["filter_shape"] = gf.FilterShape[filter_args["filter_shape"]]
    flt = gf.Filter(grid_type=gt, grid_vars=grid_vars, **filter_args)
    planes = [grid_vars[n] for n in flt.Laplacian.required_grid_args()]
    if issubclass(flt.Laplacian, kernels.BaseVectorLaplacian):
        fn = fmod._create_filter_func_vec(flt.filter_spec, flt.Laplacian)
        return fn(fields[0], fields[1], *planes), flt
    fn = fmod._create_filter_func(flt.filter_spec, flt.Laplacian)
    return fn(fields[0], *planes), flt
