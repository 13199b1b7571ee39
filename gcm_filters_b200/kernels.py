"""Grid-aware discrete Laplacians -- host side.

Mirrors the operator protocol of the reference (``gcm_filters/kernels.py``): a class per
:class:`GridType` registered in :data:`ALL_KERNELS`, constructed as ``Laplacian(**grid_vars)``,
exposing ``prepare`` / ``__call__`` / ``finalize``, the classmethod ``required_grid_args()`` and the
class attribute ``is_dimensional``.  What differs is *where the work happens*: ``__post_init__``
validates the grid variables (same exceptions and messages as the reference) and **precombines**
them once, on the host, into the few coefficient planes the CUDA stencil families of ``libgcmf.so``
read (``include/gcmf.h``); ``__call__`` and the filter loop then run entirely on the GPU.  The
reference recomputes its derived arrays on every call / dask block (``filter.py:183``).

There is no numpy implementation of the Laplacians here and no CPU fallback.
"""
import enum

import numpy as np

from . import _cabi
from . import engine

# same members, same order as the reference (kernels.py:13-28)
GridType = enum.Enum(
    "GridType",
    [
        "REGULAR",
        "REGULAR_AREA_WEIGHTED",
        "REGULAR_WITH_LAND",
        "REGULAR_WITH_LAND_AREA_WEIGHTED",
        "IRREGULAR_WITH_LAND",
        "MOM5U",
        "MOM5T",
        "TRIPOLAR_REGULAR_WITH_LAND_AREA_WEIGHTED",
        "TRIPOLAR_POP_WITH_LAND",
        "VECTOR_C_GRID",
        "VECTOR_B_GRID",
    ],
)

ALL_KERNELS = {}


# ---------------------------------------------------------------------------------------------
# small host helpers
# ---------------------------------------------------------------------------------------------
def _host(a):
    """Grid variable -> numpy array (accepts numpy, torch tensors, xarray DataArrays, lists)."""
    if hasattr(a, "detach") and hasattr(a, "cpu"):  # torch tensor
        return a.detach().cpu().numpy()
    if hasattr(a, "values") and hasattr(a, "dims"):  # xarray.DataArray
        return np.asarray(a.values)
    return np.asarray(a)


def _shift(a, dj, di):
    """value at [j+dj, i+di] with periodic wrap on the last two axes."""
    out = a
    if dj:
        out = np.roll(out, -dj, axis=-2)
    if di:
        out = np.roll(out, -di, axis=-1)
    return out


def _east(a):
    return _shift(a, 0, 1)


def _west(a):
    return _shift(a, 0, -1)


def _north(a):
    return _shift(a, 1, 0)


def _south(a):
    return _shift(a, -1, 0)


def _wet(mask):
    """any numeric / bool mask -> uint8 (value != 0 is ocean).  The reference needs float or signed
    masks (unsigned ones silently break its ``-wet_fac``, bool raises); here any dtype works."""
    return (_host(mask) != 0).astype(np.uint8)


def _f64(a):
    return _host(a).astype(np.float64, copy=False)


class Planes:
    """Result of the host precombination: what the device operator needs."""

    def __init__(self, op, flags, planes, mask=None):
        self.op = op
        self.flags = flags
        self.planes = planes  # list of float64 arrays (..., ny, nx) or None per slot
        self.mask = mask      # uint8 array for slot 0 of OP_REGULAR5, else None


# ---------------------------------------------------------------------------------------------
# base classes (kernels.py:43-104)
# ---------------------------------------------------------------------------------------------
class _BaseLaplacian:
    is_dimensional = False
    ncomp = 1
    _grid_args = []

    def __init__(self, *args, **grid_vars):
        names = self.required_grid_args()
        if args:
            if len(args) > len(names):
                raise TypeError(f"{type(self).__name__} takes {len(names)} grid variables")
            grid_vars = dict(zip(names, args), **grid_vars)
        missing = [n for n in names if n not in grid_vars]
        extra = [n for n in grid_vars if n not in names]
        if missing or extra:
            raise TypeError(f"{type(self).__name__}() expects grid variables {names}; "
                            f"missing {missing}, unexpected {extra}")
        for n in names:
            setattr(self, n, grid_vars[n])
        self._float_dtypes = [_host(grid_vars[n]).dtype for n in names
                              if "mask" not in n and _host(grid_vars[n]).dtype.kind == "f"]
        self._device_state = {}
        self.__post_init__()

    def __post_init__(self):
        self._planes = self._precombine()

    @classmethod
    def required_grid_args(cls):
        """Names of the grid variables, in the positional order of ``filter_func(field, *args)``
        (reference: own-class ``__annotations__`` order, kernels.py:58-63)."""
        return list(cls._grid_args)

    # the compute dtype of a call: fp32 only if the field and every floating grid variable are fp32
    def compute_dtype(self, field_dtype):
        if np.dtype(field_dtype) == np.float32 and all(d == np.float32 for d in self._float_dtypes):
            return np.dtype(np.float32)
        return np.dtype(np.float64)

    def _precombine(self):  # pragma: no cover - abstract
        raise NotImplementedError


class BaseScalarLaplacian(_BaseLaplacian):
    """Base class for scalar Laplacians (kernels.py:43-63)."""

    ncomp = 1

    def prepare(self, field):
        return field

    def __call__(self, field):
        """One application of the Laplacian on the GPU; numpy in -> numpy out, torch in -> torch out."""
        return engine.run_laplacian(self, (field,))[0]

    def finalize(self, field):
        return field


class BaseVectorLaplacian(_BaseLaplacian):
    """Base class for vector Laplacians (kernels.py:66-86)."""

    ncomp = 2

    def prepare(self, ufield, vfield):
        return (ufield, vfield)

    def __call__(self, ufield, vfield):
        return engine.run_laplacian(self, (ufield, vfield))

    def finalize(self, ufield, vfield):
        return (ufield, vfield)


class AreaWeightedMixin:
    """Weight and de-weight a field by the cell area (kernels.py:89-104).  Inside the filter both
    happen in the filter kernels; these methods serve direct callers of the operator protocol and run
    on the device as well (gcmf_prepare / gcmf_finalize)."""

    def prepare(self, field):
        return engine.run_area_op(self, field, divide=False)

    def finalize(self, field):
        return engine.run_area_op(self, field, divide=True)


# ---------------------------------------------------------------------------------------------
# scalar Laplacians on the unit grid  (device family GCMF_OP_REGULAR5)
# ---------------------------------------------------------------------------------------------
class RegularLaplacian(BaseScalarLaplacian):
    """Scalar Laplacian for regularly spaced Cartesian grids (kernels.py:107-124)."""

    is_dimensional = False
    _grid_args = []

    def _precombine(self):
        return Planes(_cabi.OP_REGULAR5, _cabi.FLAG_WRAP_Y, [])


ALL_KERNELS[GridType.REGULAR] = RegularLaplacian


class RegularLaplacianWithArea(AreaWeightedMixin, RegularLaplacian):
    """Regular Laplacian on the area-weighted field (kernels.py:127-147)."""

    is_dimensional = False
    _grid_args = ["area"]

    def _precombine(self):
        return Planes(_cabi.OP_REGULAR5, _cabi.FLAG_WRAP_Y | _cabi.FLAG_AREA, [None, _f64(self.area)])


ALL_KERNELS[GridType.REGULAR_AREA_WEIGHTED] = RegularLaplacianWithArea


class RegularLaplacianWithLandMask(BaseScalarLaplacian):
    """Regular Laplacian with a land mask (kernels.py:150-190).  The device kernel recomputes
    ``wet_fac`` (:165-170) from the four neighbouring uint8 mask bytes."""

    is_dimensional = False
    _grid_args = ["wet_mask"]

    def _precombine(self):
        return Planes(_cabi.OP_REGULAR5, _cabi.FLAG_WRAP_Y | _cabi.FLAG_MASK | _cabi.FLAG_NAN2NUM, [None],
                      mask=_wet(self.wet_mask))


ALL_KERNELS[GridType.REGULAR_WITH_LAND] = RegularLaplacianWithLandMask


class RegularLaplacianWithLandMaskAndArea(AreaWeightedMixin, RegularLaplacianWithLandMask):
    """kernels.py:193-219."""

    is_dimensional = False
    _grid_args = ["area", "wet_mask"]

    def _precombine(self):
        return Planes(_cabi.OP_REGULAR5,
                      _cabi.FLAG_WRAP_Y | _cabi.FLAG_MASK | _cabi.FLAG_NAN2NUM | _cabi.FLAG_AREA,
                      [None, _f64(self.area)], mask=_wet(self.wet_mask))


ALL_KERNELS[GridType.REGULAR_WITH_LAND_AREA_WEIGHTED] = RegularLaplacianWithLandMaskAndArea


def _require_land_south_row(mask_u8):
    # kernels.py:457-459, 520-522
    if mask_u8[..., 0, :].any():
        raise AssertionError("Wet mask requires zeros in southernmost row")


class TripolarRegularLaplacianTpoint(AreaWeightedMixin, BaseScalarLaplacian):
    """Area-weighted regular Laplacian with land mask and tripolar fold (kernels.py:435-492).
    The reference appends a mirrored row to every array on every call (:33-40, :474); the device
    kernel instead takes the north neighbour of (ny-1, i) from (ny-1, nx-1-i)."""

    is_dimensional = False
    _grid_args = ["area", "wet_mask"]

    def _precombine(self):
        m = _wet(self.wet_mask)
        _require_land_south_row(m)
        flags = (_cabi.FLAG_WRAP_Y | _cabi.FLAG_MASK | _cabi.FLAG_NAN2NUM | _cabi.FLAG_AREA
                 | _cabi.FLAG_FOLD_N | _cabi.FLAG_CUT_S)
        return Planes(_cabi.OP_REGULAR5, flags, [None, _f64(self.area)], mask=m)


ALL_KERNELS[GridType.TRIPOLAR_REGULAR_WITH_LAND_AREA_WEIGHTED] = TripolarRegularLaplacianTpoint


# ---------------------------------------------------------------------------------------------
# flux-form scalar Laplacians  (device family GCMF_OP_FLUX):
#   Lap[j,i] = (((Fe[j,i] - Fe[j,i-1]) + Fn[j,i]) - Fn[j-1,i]) * ra[j,i]
#   Fe[j,i] = (o[j,i+1]-o[j,i]) * ce[j,i],  Fn[j,i] = (o[j+1,i]-o[j,i]) * cn[j,i],  o = nan_to_num(f)
# The host folds spacings, edge masks and kappas into ce / cn / ra (SURVEY.md note N3: agrees with
# the reference's division form to ~1e-16 relative).  Faces next to land get an exact 0.
# ---------------------------------------------------------------------------------------------
def _face(coef, open_face):
    """coefficient on open faces, exact 0 on closed ones (also where the metric is inf/NaN on land)."""
    return np.where(open_face != 0, coef, 0.0)


_FLUX_FLAGS = _cabi.FLAG_WRAP_Y | _cabi.FLAG_NAN2NUM


class IrregularLaplacianWithLandMask(BaseScalarLaplacian):
    """Scalar Laplacian for locally orthogonal grids with land mask and spatially varying
    nondimensional diffusivities kappa_w / kappa_s (kernels.py:222-318)."""

    is_dimensional = True
    _grid_args = ["wet_mask", "dxw", "dyw", "dxs", "dys", "area", "kappa_w", "kappa_s"]

    def _precombine(self):
        kw, ks = _f64(self.kappa_w), _f64(self.kappa_s)
        if np.any(kw > 1.0):  # kernels.py:262-266
            raise ValueError("There are kappa_w values > 1 and this can cause the filter to blow up."
                             "Please make sure all kappa_w are <=1.")
        if np.any(ks > 1.0):  # kernels.py:268-272
            raise ValueError("There are kappa_s values > 1 and this can cause the filter to blow up."
                             "Please make sure all kappa_s are <=1.")
        if not (np.any(np.isclose(kw, 1.0, rtol=0, atol=1e-05)) or np.any(np.isclose(ks, 1.0, rtol=0, atol=1e-05))):
            raise ValueError(  # kernels.py:274-281
                "At least one place in the domain must have either kappa_w = 1 or kappa_s = 1. "
                "Otherwise the filter's scale will not be equal to filter_scale anywhere in the domain.")
        m = _wet(self.wet_mask).astype(np.float64)
        with np.errstate(divide="ignore", invalid="ignore"):
            w_open = m * _west(m)   # west face of (j,i) is open      (kernels.py:286-288)
            s_open = m * _south(m)  # south face                      (kernels.py:293-295)
            cw = _face(_f64(self.dyw) / _f64(self.dxw) * (w_open * kw), w_open)   # :302-304, :309
            cs = _face(_f64(self.dxs) / _f64(self.dys) * (s_open * ks), s_open)   # :305-307, :310
            ra = 1.0 / _f64(self.area)                                            # :314
        # east face of (j,i) is the west face of (j,i+1); north face is the south face of (j+1,i)
        return Planes(_cabi.OP_FLUX, _FLUX_FLAGS, [_east(cw), _north(cs), ra])


ALL_KERNELS[GridType.IRREGULAR_WITH_LAND] = IrregularLaplacianWithLandMask


class MOM5LaplacianU(BaseScalarLaplacian):
    """Laplacian for MOM5 velocity points (kernels.py:321-375).  The index pattern of the reference
    is reproduced as coded: the north-face flux is masked with wet[j,i]*wet[j,i+1] and the
    east-face flux with wet[j,i]*wet[j+1,i] (:348-349, :358-359)."""

    is_dimensional = True
    _grid_args = ["wet_mask", "dxt", "dyt", "dxu", "dyu", "area_u"]

    def _precombine(self):
        m = _wet(self.wet_mask).astype(np.float64)
        dxt, dyt, dxu, dyu = _f64(self.dxt), _f64(self.dyt), _f64(self.dxu), _f64(self.dyu)
        xm, ym = m * _east(m), m * _north(m)
        with np.errstate(divide="ignore", invalid="ignore"):
            cn = 2.0 / (_north(dxt) + _shift(dxt, 1, 1)) * (0.5 * (dyu + _north(dyu)))  # :354-355, :361
            ce = 2.0 / (_east(dyt) + _shift(dyt, 1, 1)) * (0.5 * (dxu + _east(dxu)))    # :356-357, :367
            ra = 1.0 / _f64(self.area_u)
        return Planes(_cabi.OP_FLUX, _FLUX_FLAGS, [_face(ce, ym), _face(cn, xm), ra])


ALL_KERNELS[GridType.MOM5U] = MOM5LaplacianU


class MOM5LaplacianT(BaseScalarLaplacian):
    """Laplacian for MOM5 tracer points (kernels.py:378-432); masks as coded in the reference."""

    is_dimensional = True
    _grid_args = ["wet_mask", "dxt", "dyt", "dxu", "dyu", "area_t"]

    def _precombine(self):
        m = _wet(self.wet_mask).astype(np.float64)
        dxt, dyt, dxu, dyu = _f64(self.dxt), _f64(self.dyt), _f64(self.dxu), _f64(self.dyu)
        xm, ym = m * _east(m), m * _north(m)
        with np.errstate(divide="ignore", invalid="ignore"):
            cn = 2.0 / (dxu + _west(dxu)) * (0.5 * (dyt + _north(dyt)))   # :411-412, :418
            ce = 2.0 / (dyu + _south(dyu)) * (0.5 * (dxt + _east(dxt)))   # :413-414, :424
            ra = 1.0 / _f64(self.area_t)
        return Planes(_cabi.OP_FLUX, _FLUX_FLAGS, [_face(ce, ym), _face(cn, xm), ra])


ALL_KERNELS[GridType.MOM5T] = MOM5LaplacianT


class POPTripolarLaplacianTpoint(BaseScalarLaplacian):
    """Scalar Laplacian for POP's tripolar grid, T points (kernels.py:495-588)."""

    is_dimensional = True
    _grid_args = ["wet_mask", "dxe", "dye", "dxn", "dyn", "tarea"]

    def _precombine(self):
        mu8 = _wet(self.wet_mask)
        _require_land_south_row(mu8)
        m = mu8.astype(np.float64)
        dxn, dyn = _f64(self.dxn), _f64(self.dyn)
        e_open = m * _east(m)                                   # :538
        n_open = m * _north(m)                                  # :543 for rows < ny-1 ...
        n_open[..., -1, :] = m[..., -1, :] * m[..., -1, ::-1]   # ... and across the fold for row ny-1 (:33-40)
        # the northernmost row of dxn, dyn must fold onto itself where the fold face is open (:545-562)
        half = dxn.shape[-1] // 2
        row = np.where(n_open[..., -1, :] == 1, dxn[..., -1, :], 0)
        if not np.all(row[..., :half][..., ::-1] == row[..., half:]):
            raise AssertionError("Northernmost row of dxn does not fold onto itself. "
                                 "This is a requirement for using a tripole boundary condition.")
        row = np.where(n_open[..., -1, :] == 1, dyn[..., -1, :], 0)
        if not np.allclose(row[..., :half][..., ::-1], row[..., half:]):
            raise AssertionError("Northernmost row of dyn does not fold onto itself. "
                                 "This is a requirement for using a tripole boundary condition.")
        with np.errstate(divide="ignore", invalid="ignore"):
            ce = _face(_f64(self.dye) / _f64(self.dxe), e_open)  # :571-573, :578
            cn = _face(dxn / dyn, n_open)                        # :574-576, :579
            ra = 1.0 / _f64(self.tarea)                          # :584
        flags = _FLUX_FLAGS | _cabi.FLAG_FOLD_N | _cabi.FLAG_CUT_S
        return Planes(_cabi.OP_FLUX, flags, [ce, cn, ra])


ALL_KERNELS[GridType.TRIPOLAR_POP_WITH_LAND] = POPTripolarLaplacianTpoint


# ---------------------------------------------------------------------------------------------
# vector Laplacians
# ---------------------------------------------------------------------------------------------
class CgridVectorLaplacian(BaseVectorLaplacian):
    """Vector Laplacian on a C-grid after Griffies & Hallberg 2000 (kernels.py:591-699).
    The 14 grid variables are folded into the 14 planes listed in include/gcmf.h."""

    is_dimensional = True
    _grid_args = ["wet_mask_t", "wet_mask_q", "dxT", "dyT", "dxCu", "dyCu", "dxCv", "dyCv", "dxBu", "dyBu",
                  "area_u", "area_v", "kappa_iso", "kappa_aniso"]

    def _precombine(self):
        mt = _wet(self.wet_mask_t).astype(np.float64)
        mq = _wet(self.wet_mask_q).astype(np.float64)
        dxT, dyT, dxBu, dyBu = _f64(self.dxT), _f64(self.dyT), _f64(self.dxBu), _f64(self.dyBu)
        kiso, kaniso = _f64(self.kappa_iso), _f64(self.kappa_aniso)
        au, av = _f64(self.area_u), _f64(self.area_v)
        with np.errstate(divide="ignore", invalid="ignore"):
            kt = kiso + 0.5 * kaniso                                  # :661
            planes = [
                1.0 / _f64(self.dyCu), 1.0 / _f64(self.dxCv), 1.0 / _f64(self.dyCv), 1.0 / _f64(self.dxCu),
                _face(kt * (dyT / dxT), mt), _face(kt * (dxT / dyT), mt),        # :633-634, :661
                _face(kiso * (dyBu / dxBu), mq), _face(kiso * (dxBu / dyBu), mq),  # :635-636, :670
                dyT * dyT, dxT * dxT, dxBu * dxBu, dyBu * dyBu,                # :638-641
                np.where(au > 0, 1.0 / au, 0.0), np.where(av > 0, 1.0 / av, 0.0),  # :644-645
            ]
        shape = np.broadcast_shapes(*[p.shape for p in planes])
        planes = [np.ascontiguousarray(np.broadcast_to(p, shape)) for p in planes]
        return Planes(_cabi.OP_VECTOR_C, _FLUX_FLAGS, planes)


ALL_KERNELS[GridType.VECTOR_C_GRID] = CgridVectorLaplacian


class BgridVectorLaplacian(BaseVectorLaplacian):
    """Vector Laplacian on a B-grid, POP formulation, periodic (kernels.py:702-840).  The ten stencil
    coefficients depend only on the grid; the reference rebuilds them on every call (:751-805), here
    they are built once, in the same arithmetic order, and the kernel does the 10-term sums."""

    is_dimensional = True
    _grid_args = ["DXU", "DYU", "HUS", "HUW", "HTE", "HTN", "UAREA", "TAREA"]

    def _precombine(self):
        DXU, DYU, HUS, HUW = _f64(self.DXU), _f64(self.DYU), _f64(self.HUS), _f64(self.HUW)
        HTE, HTN = _f64(self.HTE), _f64(self.HTN)
        uar, tar = 1 / _f64(self.UAREA), 1 / _f64(self.TAREA)
        dxur, dyur = 1 / DXU, 1 / DYU
        r_se = HUS / HTE
        dus, dun = r_se * uar, _west(r_se) * uar
        r_wn = HUW / HTN
        duw, due = r_wn * uar, _south(r_wn) * uar
        kxu = (_south(HUW) - HUW) * uar
        kyu = (_west(HUS) - HUS) * uar
        kxt = (HTE - _north(HTE)) * tar
        avg = 0.5 * (kxt + _west(kxt))
        dxkx = (_south(avg) - avg) * dxur
        avg = 0.5 * (kxt + _south(kxt))
        dykx = (_west(avg) - avg) * dyur
        kyt = (HTN - _east(HTN)) * tar
        avg = 0.5 * (kyt + _south(kyt))
        dyky = (_west(avg) - avg) * dyur
        avg = 0.5 * (kyt + _west(kyt))
        dxky = (_south(avg) - avg) * dxur
        dum = -(dxkx + dyky + 2 * (kxu * kxu + kyu * kyu))
        dmc = dxky - dykx
        dme = (2 * kyu) / (HTN + _south(HTN))
        dmn = -(2 * kxu) / (HTE + _west(HTE))
        duc = -(dun + dus + due + duw)
        cc = duc + dum
        planes = [cc, dun, dus, due, duw, dmc, dmn, dme]
        return Planes(_cabi.OP_VECTOR_B, _FLUX_FLAGS, planes)


ALL_KERNELS[GridType.VECTOR_B_GRID] = BgridVectorLaplacian


def required_grid_vars(grid_type):
    """Names of the grid variables a grid type needs (kernels.py:843-858)."""
    return ALL_KERNELS[grid_type].required_grid_args()
