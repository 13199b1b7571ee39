"""Main Filter class -- same constructor, attributes, errors and warnings as the reference
(``gcm_filters/filter.py``), with the Chebyshev step loop executed by libgcmf.so on a B200.

Host-side work kept in Python (it is O(n_steps^2) scalar work, done once per Filter):
``_compute_n_steps_default`` and ``_compute_filter_spec`` (reference filter.py:74-151).
"""
import enum
import warnings
from dataclasses import dataclass, field
from typing import Iterable, NamedTuple

import numpy as np

from . import engine
from .kernels import ALL_KERNELS, AreaWeightedMixin, BaseScalarLaplacian, BaseVectorLaplacian, GridType

FilterShape = enum.Enum("FilterShape", ["GAUSSIAN", "TAPER"])

# n_steps_default = ceil((offset + factor*(pi/transition_width)**exponent) * filter_scale/dx_min)
# (reference filter.py:28-37)
filter_params = {
    FilterShape.GAUSSIAN: {1: {"offset": 0.8, "factor": 0.0, "exponent": 1},
                           2: {"offset": 1.1, "factor": 0.0, "exponent": 1}},
    FilterShape.TAPER: {1: {"offset": 2.2, "factor": 0.6, "exponent": 2.5},
                        2: {"offset": 3.2, "factor": 0.7, "exponent": 2.7}},
}


class TargetSpec(NamedTuple):
    s_max: float
    filter_scale: float
    transition_width: float


def _gaussian_target(ts: TargetSpec):
    """exp(-s L^2/24) with s = s_max (t+1)/2 (reference filter.py:47-50)."""
    return lambda t: np.exp(-(ts.s_max * (t + 1) / 2) * (ts.filter_scale) ** 2 / 24)


def _taper_target(ts: TargetSpec):
    """Piecewise-cubic taper through (0,1), (2pi/(w L),1), (2pi/L,0), (8 sqrt(s_max),0)
    evaluated at k = sqrt(s) (reference filter.py:53-65)."""
    from scipy import interpolate

    knots = np.array([0, 2 * np.pi / (ts.transition_width * ts.filter_scale), 2 * np.pi / ts.filter_scale,
                      8 * np.sqrt(ts.s_max)])
    fk = interpolate.PchipInterpolator(knots, np.array([1, 1, 0, 0]))
    return lambda t: fk(np.sqrt((t + 1) * (ts.s_max / 2)))


_target_function = {FilterShape.GAUSSIAN: _gaussian_target, FilterShape.TAPER: _taper_target}


def _compute_n_steps_default(ndim, filter_shape, filter_scale, dx_min, transition_width):
    """Default number of steps for 1-D / 2-D filters (reference filter.py:74-89)."""
    prm = filter_params[filter_shape][ndim]
    n_steps_factor = prm["offset"] + prm["factor"] * ((np.pi / transition_width) ** prm["exponent"])
    return max(np.ceil(n_steps_factor * (filter_scale / dx_min)).astype(int), 3)


class FilterSpec(NamedTuple):
    n_steps: int
    s_max: float
    p: Iterable[float]
    dx_min_sq: float


def _compute_filter_spec(filter_scale, dx_min, filter_shape, transition_width=np.pi, ndim=2, n_steps=0):
    """Chebyshev coefficients p[0..n_steps] of the polynomial approximating the target filter:
    Galerkin projection in Shen's basis phi_i = T_i - T_{i+2} with Chebyshev-Gauss quadrature
    (reference filter.py:99-151; same operation order so the coefficients agree to the last bit)."""
    n = n_steps
    mass = (np.pi / 2) * (2 * np.eye(n - 1) - np.diag(np.ones(n - 3), 2) - np.diag(np.ones(n - 3), -2))
    mass[0, 0] = 3 * np.pi / 2
    s_max = ndim * (2 / dx_min) ** 2
    F = _target_function[filter_shape](TargetSpec(s_max, filter_scale, transition_width))
    nodes, weights = np.polynomial.chebyshev.chebgauss(n + 1)
    mismatch = F(nodes) - ((1 - nodes) / 2 + F(1) * (nodes + 1) / 2)
    rhs = np.zeros(n - 1)
    for i in range(n - 1):
        basis = np.zeros(n + 1)
        basis[i] = 1
        basis[i + 2] = -1
        rhs[i] = np.sum(weights * np.polynomial.chebyshev.chebval(nodes, basis) * mismatch)
    c_hat = np.linalg.solve(mass, rhs)
    p = np.zeros(n + 1)
    p[0] = c_hat[0] + (1 + F(1)) / 2
    p[1] = c_hat[1] - (1 - F(1)) / 2
    p[2:n - 1] = c_hat[2:n - 1] - c_hat[0:n - 3]
    p[n - 1] = -c_hat[n - 3]
    p[n] = -c_hat[n - 2]
    return FilterSpec(n, s_max, p, dx_min ** 2)


def _shift_scale(filter_spec: FilterSpec, Laplacian):
    """c in A(x) = -x - c*Lap(x) (reference filter.py:168-173)."""
    if Laplacian.is_dimensional:
        return 2 / filter_spec.s_max
    return 2 / (filter_spec.s_max * filter_spec.dx_min_sq)


_FP_WHOLE_BYTES = 4 << 20
_FP_ROWS = 64


def _fingerprint(a):
    """Cheap content fingerprint of a host array: the wrap-around sum of its 8-byte words.  Arrays up to 4 MiB are
    summed whole (any in-place edit shows); larger ones are sampled on 64 evenly spaced rows, first and last included
    (~0.3 ms for a 69 MB plane -- the check runs on every filter call).  Call ``Filter.invalidate_grid_cache()`` after
    editing a large grid variable in place if the edit may fall between the sampled rows."""
    if a.nbytes > _FP_WHOLE_BYTES and a.ndim >= 2:
        rows = a.reshape((-1, a.shape[-1])) if a.flags.c_contiguous else a.reshape((-1, a.shape[-1]))
        idx = np.unique(np.linspace(0, rows.shape[0] - 1, _FP_ROWS).astype(np.int64))
        a = rows[idx]
    a = np.ascontiguousarray(a)
    raw = a.reshape(-1).view(np.uint8)
    n8 = raw.shape[0] // 8 * 8
    total = int(raw[:n8].view(np.uint64).sum(dtype=np.uint64)) if n8 else 0
    return (total + int(raw[n8:].sum(dtype=np.uint64))) & 0xFFFFFFFFFFFFFFFF


class _LaplacianCache:
    """Laplacian objects (validated + precombined + uploaded planes) so that repeated calls / dask blocks do not
    rebuild them as the reference does on every call (filter.py:183).  Keyed by the identity AND a content fingerprint
    of the grid arrays: an in-place edit of a grid variable between two calls rebuilds the operator, as the reference
    would.  The cache keeps the (converted) arrays alive, so an address cannot be recycled by a different array."""

    def __init__(self, Laplacian, size=4):
        self.Laplacian = Laplacian
        self.size = size
        self.entries = []

    @staticmethod
    def _key(a):
        """(key, array to keep alive)"""
        if engine._is_torch(a):
            return ("t", a.data_ptr(), tuple(a.shape), str(a.dtype), tuple(a.stride()), a._version), a
        arr = np.asarray(getattr(a, "values", a))
        return ("n", arr.__array_interface__["data"][0], arr.shape, arr.dtype.str, arr.strides, _fingerprint(arr)), arr

    def clear(self):
        self.entries = []

    def get(self, args):
        keyed = [self._key(a) for a in args]
        key = tuple(k for k, _ in keyed)
        for k, keep, lap in self.entries:
            if k == key:
                return lap
        names = self.Laplacian.required_grid_args()
        lap = self.Laplacian(**dict(zip(names, args)))
        self.entries.append((key, (args, [arr for _, arr in keyed]), lap))
        if len(self.entries) > self.size:
            self.entries.pop(0)
        return lap


def _create_filter_func(filter_spec: FilterSpec, Laplacian: BaseScalarLaplacian, _cache=None):
    """Returns ``filter_func(field, *grid_args)`` with the reference's signature (filter.py:154-214):
    arrays arrive with the two filtered dims last (y then x), batch dims leading."""
    cache = _cache or _LaplacianCache(Laplacian)
    c = _shift_scale(filter_spec, Laplacian)

    def filter_func(field, *args, out=None):
        assert len(args) == len(Laplacian.required_grid_args())
        laplacian = cache.get(args)
        return engine.run_filter(laplacian, filter_spec.p, c, (field,), out=out)[0]

    return filter_func


def _create_filter_func_vec(filter_spec: FilterSpec, Laplacian: BaseVectorLaplacian, _cache=None):
    """Returns ``filter_func_vec(ufield, vfield, *grid_args)`` (reference filter.py:217-291)."""
    cache = _cache or _LaplacianCache(Laplacian)
    c = _shift_scale(filter_spec, Laplacian)

    def filter_func_vec(ufield, vfield, *args, out=None):
        assert len(args) == len(Laplacian.required_grid_args())
        laplacian = cache.get(args)
        return engine.run_filter(laplacian, filter_spec.p, c, (ufield, vfield), out=out)

    return filter_func_vec


def _xarray():
    try:
        import xarray as xr

        return xr if hasattr(xr, "apply_ufunc") else None
    except ImportError:
        return None


@dataclass
class Filter:
    """A class for applying diffusion-based smoothing filters to gridded data.

    Parameters (identical to the reference, filter.py:294-333)
    ----------
    filter_scale : float
    dx_min : float
    filter_shape : FilterShape
    transition_width : float, optional
    ndim : int, optional
    n_steps : int, optional (0 = choose automatically)
    grid_type : GridType
    grid_vars : dict of grid variables (numpy arrays, torch tensors or xarray DataArrays)

    Attributes
    ----------
    filter_spec: FilterSpec
    """

    filter_scale: float
    dx_min: float
    filter_shape: FilterShape = FilterShape.GAUSSIAN
    transition_width: float = np.pi
    ndim: int = 2
    n_steps: int = 0
    grid_type: GridType = GridType.REGULAR
    grid_vars: dict = field(default_factory=dict, repr=False)

    def __post_init__(self):
        self.Laplacian = ALL_KERNELS[self.grid_type]

        # simple fixed factor filters work on a transformed grid with dx = dy = 1 (filter.py:339-346)
        if issubclass(self.Laplacian, AreaWeightedMixin):
            if self.dx_min != 1:
                raise ValueError(
                    "Provided Laplacian is for simple fixed factor filtering, "
                    "where transformed field is filtered on a regular grid with dx = dy = 1. "
                    "dx_min must be set to 1."
                )
        if self.transition_width <= 1:  # filter.py:349-350
            raise ValueError("Transition width must be > 1.")
        if self.ndim > 2:  # filter.py:353-357
            if self.n_steps < 3:
                raise ValueError("When ndim > 2, you must set n_steps manually")
            n_steps_default = self.n_steps
        else:
            n_steps_default = _compute_n_steps_default(
                self.ndim, self.filter_shape, self.filter_scale, self.dx_min, self.transition_width)
        if self.n_steps < 3:  # filter.py:368-369
            self.n_steps = n_steps_default
        if self.n_steps < n_steps_default:  # filter.py:371-375
            warnings.warn("You have set n_steps below the default. Results might not be accurate.", stacklevel=2)

        self.filter_spec = _compute_filter_spec(
            self.filter_scale, self.dx_min, self.filter_shape, self.transition_width, self.ndim, self.n_steps)

        if not set(self.Laplacian.required_grid_args()) == set(self.grid_vars):  # filter.py:388-392
            raise ValueError(
                f"Provided `grid_vars` {list(self.grid_vars)} do not match expected "
                f"{list(self.Laplacian.required_grid_args())}"
            )
        xr = _xarray()
        if xr is not None and all(isinstance(v, xr.DataArray) for v in self.grid_vars.values()) and self.grid_vars:
            self.grid_ds = xr.Dataset({name: da for name, da in self.grid_vars.items()})  # filter.py:393
        else:
            self.grid_ds = dict(self.grid_vars)
        self._cache = _LaplacianCache(self.Laplacian)

    # --------------------------------------------------------------------------------------
    @property
    def laplacian(self):
        """The Laplacian operator built from this filter's own grid variables (cached)."""
        args = [self.grid_ds[name] for name in self.Laplacian.required_grid_args()]
        return self._cache.get(tuple(args))

    def invalidate_grid_cache(self):
        """Forget the cached Laplacian operators (validated, precombined and uploaded grid variables).  Only needed
        after editing a LARGE grid variable in place between two calls (see ``_fingerprint``); the reference rebuilds
        its Laplacian on every call (filter.py:183)."""
        self._cache.clear()

    def plot_shape(self, ax=None):
        """Not part of the filter path (reference filter.py:395-428, matplotlib only): out of scope here, see
        DESIGN.md section 8.  ``filter_spec.p`` holds the Chebyshev coefficients a plot would evaluate."""
        raise NotImplementedError("plot_shape is outside the scope of gcm_filters_b200 (the iterative filter path); "
                                  "use the reference package to plot the target filter")

    # --------------------------------------------------------------------------------------
    def apply(self, ds, dims=None, out=None):
        """Filter a field with a scalar Laplacian across ``dims``.

        ``ds`` may be an ``xarray.DataArray`` / ``xarray.Dataset`` (when xarray is installed; same
        semantics as the reference, filter.py:430-469, dimension order matters: y first), or a
        plain numpy array / torch tensor whose LAST two axes are (y, x); leading axes are batch
        dimensions.  For arrays ``dims`` is only checked for length.  A ``dict`` of such arrays (several
        variables on one grid) returns a dict; numpy variables of one dtype share one batched call."""
        if issubclass(self.Laplacian, BaseVectorLaplacian):
            raise ValueError(
                f"Provided Laplacian {self.Laplacian} is a vector Laplacian. "
                f"The ``.apply`` method is only suitable for scalar Laplacians."
            )
        xr = _xarray()
        if xr is not None and isinstance(ds, xr.Dataset):
            filtered = ds.copy(deep=True)
            any_filtered = False
            for key, var in filtered.variables.items():
                if all(dim in var.dims for dim in dims):
                    filtered[key] = self._apply_to_dataarray(var, dims=dims)
                    any_filtered = True
            if not any_filtered:
                warnings.warn(
                    f"No variables in the dataset had all of the given "
                    f"dimensions ({dims}), so nothing was filtered.",
                    stacklevel=2,
                )
            return filtered
        if isinstance(ds, dict):
            return self._apply_to_mapping(ds, dims)
        return self._apply_to_dataarray(ds, dims=dims, out=out)

    def _apply_to_mapping(self, variables, dims):
        """``{name: array}`` of variables that live on the same horizontal grid (the array counterpart of the reference's
        Dataset loop, filter.py:454-467): numpy arrays of one dtype are filtered in ONE batched call -- their batch
        axes are flattened and concatenated, so the coefficient planes are staged once and the launches are as long
        as for a single large variable -- anything else (device tensors, mixed dtypes) variable by variable."""
        if not variables:
            return {}
        names = list(variables)
        arrs = [variables[k] for k in names]
        plain = all(isinstance(a, np.ndarray) and a.ndim >= 2 for a in arrs)
        if plain and len({a.dtype for a in arrs}) == 1 and len({a.shape[-2:] for a in arrs}) == 1 and len(arrs) > 1:
            ny, nx = arrs[0].shape[-2:]
            flat = [a.reshape((-1, ny, nx)) for a in arrs]
            res = self._apply_to_dataarray(np.concatenate(flat), dims=dims)
            out, b0 = {}, 0
            for k, a, f in zip(names, arrs, flat):
                out[k] = res[b0:b0 + f.shape[0]].reshape(a.shape)
                b0 += f.shape[0]
            return out
        return {k: self._apply_to_dataarray(a, dims=dims) for k, a in zip(names, arrs)}

    def _grid_args(self):
        return [self.grid_ds[name] for name in self.Laplacian.required_grid_args()]

    def _apply_to_dataarray(self, field, dims, out=None):
        filter_func = _create_filter_func(self.filter_spec, self.Laplacian, self._cache)
        grid_args = self._grid_args()
        if dims is not None:
            assert len(dims) == 2
        xr = _xarray()
        if xr is not None and isinstance(field, (xr.DataArray, xr.Variable)):
            n_args = 1 + len(grid_args)
            return xr.apply_ufunc(  # reference filter.py:478-486
                filter_func, field, *grid_args,
                input_core_dims=n_args * [dims], output_core_dims=[dims],
                output_dtypes=[field.dtype], dask="parallelized",
            )
        return filter_func(field, *grid_args, out=out)

    def apply_to_vector(self, ufield, vfield, dims=None, out=None):
        """Filter a vector field with a vector Laplacian across ``dims`` (reference filter.py:490-529)."""
        if not issubclass(self.Laplacian, BaseVectorLaplacian):
            raise ValueError(
                f"Provided Laplacian {self.Laplacian} is a scalar Laplacian. "
                f"The ``.apply_to_vector`` method is only suitable for vector Laplacians."
            )
        filter_func_vec = _create_filter_func_vec(self.filter_spec, self.Laplacian, self._cache)
        grid_args = self._grid_args()
        if dims is not None:
            assert len(dims) == 2
        xr = _xarray()
        if xr is not None and isinstance(ufield, xr.DataArray):
            n_args = 2 + len(grid_args)
            return xr.apply_ufunc(
                filter_func_vec, ufield, vfield, *grid_args,
                input_core_dims=n_args * [dims], output_core_dims=2 * [dims],
                output_dtypes=[ufield.dtype, vfield.dtype], dask="parallelized",
            )
        return filter_func_vec(ufield, vfield, *grid_args, out=out)
