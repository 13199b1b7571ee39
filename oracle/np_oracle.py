"""CPU oracle: numpy restatement of the gcm-filters hot path.  TEST INFRASTRUCTURE ONLY.

Every operator follows the evaluation order of the reference (``/root/reference``,
cited per function as ``file:line``) so that the result is bit-identical to the live
reference, but is written independently in *index form*: ``E(a)[j,i] = a[j,i+1]``,
``W(a)[j,i] = a[j,i-1]``, ``N(a)[j,i] = a[j+1,i]``, ``S(a)[j,i] = a[j-1,i]`` with periodic wrap
on both axes (the reference's ``np.roll`` by -1 / +1 along axis -1 / -2).

Parity status: PINNED (see ``oracle/__init__.py``).
"""
from collections import namedtuple

import numpy as np

# ------------------------------------------------------------------ index-form shifts


def E(a):
    return np.concatenate((a[..., :, 1:], a[..., :, :1]), axis=-1)


def W(a):
    return np.concatenate((a[..., :, -1:], a[..., :, :-1]), axis=-1)


def N(a):
    return np.concatenate((a[..., 1:, :], a[..., :1, :]), axis=-2)


def S(a):
    return np.concatenate((a[..., -1:, :], a[..., :-1, :]), axis=-2)


def z(a):
    """nan_to_num: NaN -> 0, +-inf -> +-largest finite."""
    return np.nan_to_num(a)


def fold_extend(a):
    """kernels.py:33-40: append the mirrored northernmost row, (..,ny,nx) -> (..,ny+1,nx)."""
    return np.concatenate((a, a[..., -1:, ::-1]), axis=-2)


# ------------------------------------------------------------------ operator objects


class _Op:
    """Mirror of the reference operator protocol (kernels.py:43-86): prepare / apply / finalize."""

    is_dimensional = False
    ncomp = 1
    area = None  # set => AreaWeightedMixin (kernels.py:89-104)

    def prepare(self, *f):
        if self.area is None:
            return f if self.ncomp == 2 else f[0]
        return f[0] * self.area

    def finalize(self, *f):
        if self.area is None:
            return f if self.ncomp == 2 else f[0]
        return f[0] / self.area


class Regular(_Op):
    """kernels.py:107-124 (and :127-147 when ``area`` is given)."""

    def __init__(self, area=None):
        self.area = area

    def apply(self, f):
        return -4 * f + E(f) + W(f) + N(f) + S(f)


class RegularWithLand(_Op):
    """kernels.py:150-190 (and :193-219 with ``area``)."""

    def __init__(self, wet_mask, area=None):
        self.area = area
        self.m = wet_mask
        self.wet_fac = E(wet_mask) + W(wet_mask) + N(wet_mask) + S(wet_mask)  # :165-170

    def apply(self, f):
        o = self.m * z(f)  # :175-176
        o = -self.wet_fac * o + E(o) + W(o) + N(o) + S(o)  # :178-184
        return self.m * o  # :186


class IrregularWithLand(_Op):
    """kernels.py:222-318."""

    is_dimensional = True

    def __init__(self, wet_mask, dxw, dyw, dxs, dys, area, kappa_w, kappa_s):
        if np.any(kappa_w > 1.0):  # :262-266
            raise ValueError("There are kappa_w values > 1 and this can cause the filter to blow up."
                             "Please make sure all kappa_w are <=1.")
        if np.any(kappa_s > 1.0):  # :268-272
            raise ValueError("There are kappa_s values > 1 and this can cause the filter to blow up."
                             "Please make sure all kappa_s are <=1.")
        if not (np.any(np.isclose(kappa_w, 1.0, rtol=0, atol=1e-05))
                or np.any(np.isclose(kappa_s, 1.0, rtol=0, atol=1e-05))):  # :274-281
            raise ValueError("At least one place in the domain must have either kappa_w = 1 or kappa_s = 1. "
                             "Otherwise the filter's scale will not be equal to filter_scale anywhere in the domain.")
        self.dxw, self.dyw, self.dxs, self.dys, self.cell_area = dxw, dyw, dxs, dys, area
        self.wm = wet_mask * W(wet_mask) * kappa_w  # :286-288
        self.sm = wet_mask * S(wet_mask) * kappa_s  # :293-295

    def apply(self, f):
        o = z(f)  # :300
        wflux = (o - W(o)) / self.dxw * self.dyw  # :302-304
        sflux = (o - S(o)) / self.dys * self.dxs  # :305-307
        wflux = wflux * self.wm  # :309
        sflux = sflux * self.sm  # :310
        out = E(wflux) - wflux + N(sflux) - sflux  # :312
        return out / self.cell_area  # :314


class MOM5U(_Op):
    """kernels.py:321-375.  Grid variables must be 2-D (the reference rolls them on axes (0,1), :355,357)."""

    is_dimensional = True

    def __init__(self, wet_mask, dxt, dyt, dxu, dyu, area_u):
        self.dxt, self.dyt, self.dxu, self.dyu, self.area_u = dxt, dyt, dxu, dyu, area_u
        self.xm = wet_mask * E(wet_mask)  # :348
        self.ym = wet_mask * N(wet_mask)  # :349

    def apply(self, f):
        f = z(f)  # :353
        fx = 2 * (N(f) - f)  # :354
        fx = fx / (N(self.dxt) + N(E(self.dxt)))  # :355
        fy = 2 * (E(f) - f)  # :356
        fy = fy / (E(self.dyt) + N(E(self.dyt)))  # :357
        fx = fx * self.xm  # :358
        fy = fy * self.ym  # :359
        out1 = 0.5 * fx * (self.dyu + N(self.dyu))  # :361
        out1 = out1 - 0.5 * S(fx) * (self.dyu + S(self.dyu))  # :362-364
        out1 = out1 / self.area_u  # :365
        out2 = 0.5 * fy * (self.dxu + E(self.dxu))  # :367
        out2 = out2 - 0.5 * W(fy) * (self.dxu + W(self.dxu))  # :368-370
        out2 = out2 / self.area_u  # :371
        return out1 + out2  # :372


class MOM5T(_Op):
    """kernels.py:378-432."""

    is_dimensional = True

    def __init__(self, wet_mask, dxt, dyt, dxu, dyu, area_t):
        self.dxt, self.dyt, self.dxu, self.dyu, self.area_t = dxt, dyt, dxu, dyu, area_t
        self.xm = wet_mask * E(wet_mask)  # :405
        self.ym = wet_mask * N(wet_mask)  # :406

    def apply(self, f):
        f = z(f)  # :410
        fx = 2 * (N(f) - f)  # :411
        fx = fx / (self.dxu + W(self.dxu))  # :412
        fy = 2 * (E(f) - f)  # :413
        fy = fy / (self.dyu + S(self.dyu))  # :414
        fx = fx * self.xm  # :415
        fy = fy * self.ym  # :416
        out1 = fx * 0.5 * (self.dyt + N(self.dyt))  # :418
        out1 = out1 - S(fx) * 0.5 * (self.dyt + S(self.dyt))  # :419-421
        out1 = out1 / self.area_t  # :422
        out2 = fy * 0.5 * (self.dxt + E(self.dxt))  # :424
        out2 = out2 - W(fy) * 0.5 * (self.dxt + W(self.dxt))  # :425-427
        out2 = out2 / self.area_t  # :428
        return out1 + out2  # :429


class TripolarRegular(_Op):
    """kernels.py:435-492: regular 5-point on the (ny+1)-row fold-extended arrays, area weighted."""

    def __init__(self, area, wet_mask):
        if wet_mask[..., 0, :].any():  # :458-459
            raise AssertionError("Wet mask requires zeros in southernmost row")
        self.area = area
        self.m = wet_mask
        mx = fold_extend(wet_mask)  # :461
        self.wet_fac = E(mx) + W(mx) + N(mx) + S(mx)  # :462-467

    def apply(self, f):
        d = self.m * z(f)  # :472-473
        d = fold_extend(d)  # :474
        o = -self.wet_fac * d + E(d) + W(d) + N(d) + S(d)  # :476-482
        o = o[..., :-1, :]  # :484
        return self.m * o  # :486


class TripolarPOP(_Op):
    """kernels.py:495-588."""

    is_dimensional = True

    def __init__(self, wet_mask, dxe, dye, dxn, dyn, tarea):
        if wet_mask[..., 0, :].any():  # :521-522
            raise AssertionError("Wet mask requires zeros in southernmost row")
        m = fold_extend(wet_mask)  # :525
        self.dxe, self.dye = fold_extend(dxe), fold_extend(dye)  # :530-531
        self.dxn, self.dyn = fold_extend(dxn), fold_extend(dyn)  # :532-533
        self.tarea = tarea
        self.em = m * E(m)  # :538
        self.nm = m * N(m)  # :543
        nx = self.dxn.shape[-1]
        half = nx // 2
        row = np.where(self.nm == 1, self.dxn, 0)[..., -2, :]  # :549-550
        if not np.all(row[..., :half][..., ::-1] == row[..., half:]):  # :551-554
            raise AssertionError("Northernmost row of dxn does not fold onto itself. "
                                 "This is a requirement for using a tripole boundary condition.")
        row = np.where(self.nm == 1, self.dyn, 0)[..., -2, :]  # :555-556
        if not np.allclose(row[..., :half][..., ::-1], row[..., half:]):  # :559-562
            raise AssertionError("Northernmost row of dyn does not fold onto itself. "
                                 "This is a requirement for using a tripole boundary condition.")

    def apply(self, f):
        d = fold_extend(z(f))  # :566-569
        eflux = (E(d) - d) / self.dxe * self.dye  # :571-573
        nflux = (N(d) - d) / self.dyn * self.dxn  # :574-576
        eflux = eflux * self.em  # :578
        nflux = nflux * self.nm  # :579
        out = eflux - W(eflux) + nflux - S(nflux)  # :581
        out = out[..., :-1, :]  # :583
        return out / self.tarea  # :584


class VectorCGrid(_Op):
    """kernels.py:591-699."""

    is_dimensional = True
    ncomp = 2

    def __init__(self, wet_mask_t, wet_mask_q, dxT, dyT, dxCu, dyCu, dxCv, dyCv, dxBu, dyBu,
                 area_u, area_v, kappa_iso, kappa_aniso):
        self.dxCu, self.dyCu, self.dxCv, self.dyCv = dxCu, dyCu, dxCv, dyCv
        self.kappa_iso, self.kappa_aniso = kappa_iso, kappa_aniso
        self.dx_dyT = dxT / dyT * wet_mask_t  # :633
        self.dy_dxT = dyT / dxT * wet_mask_t  # :634
        self.dx_dyBu = dxBu / dyBu * wet_mask_q  # :635
        self.dy_dxBu = dyBu / dxBu * wet_mask_q  # :636
        self.dx2h, self.dy2h = dxT * dxT, dyT * dyT  # :638-639
        self.dx2q, self.dy2q = dxBu * dxBu, dyBu * dyBu  # :640-641
        with np.errstate(divide="ignore"):
            self.rau = np.where(area_u > 0, 1 / area_u, 0)  # :644
            self.rav = np.where(area_v > 0, 1 / area_v, 0)  # :645

    def apply(self, u, v):
        u, v = z(u), z(v)  # :650-651
        a = u / self.dyCu
        dudx = self.dy_dxT * (a - W(a))  # :653-655
        b = v / self.dxCv
        dvdy = self.dx_dyT * (b - S(b))  # :656-658
        sxx = dudx - dvdy  # :659
        sxx = -(self.kappa_iso + 0.5 * self.kappa_aniso) * sxx  # :661
        c = v / self.dyCv
        dvdx = self.dy_dxBu * (E(c) - c)  # :663-665
        e = u / self.dxCu
        dudy = self.dx_dyBu * (N(e) - e)  # :666-668
        sxy = dvdx + dudy  # :669
        sxy = -self.kappa_iso * sxy  # :670
        t = self.dy2h * sxx
        uc = 1 / self.dyCu * (t - E(t))  # :672-676
        q = self.dx2q * sxy
        uc = uc + 1 / self.dxCu * (S(q) - q)  # :677-681
        uc = uc * self.rau  # :682
        r = self.dy2q * sxy
        vc = 1 / self.dyCv * (W(r) - r)  # :684-688
        h = self.dx2h * sxx
        vc = vc - 1 / self.dxCv * (h - N(h))  # :689-693
        vc = vc * self.rav  # :694
        return uc, vc


class VectorBGrid(_Op):
    """kernels.py:702-840.  The stencil coefficients do not depend on the field; the reference
    recomputes them on every call (:751-805), here they are built once in the same order."""

    is_dimensional = True
    ncomp = 2

    def __init__(self, DXU, DYU, HUS, HUW, HTE, HTN, UAREA, TAREA):
        ur, tr = 1 / UAREA, 1 / TAREA  # :734-735
        dxur, dyur = 1 / DXU, 1 / DYU  # :737-738
        w1 = HUS / HTE  # :751
        self.DUS = w1 * ur  # :753
        self.DUN = W(w1) * ur  # :756
        w1 = HUW / HTN  # :760
        self.DUW = w1 * ur  # :762
        self.DUE = S(w1) * ur  # :763
        kxu = (S(HUW) - HUW) * ur  # :768-770
        kyu = (W(HUS) - HUS) * ur  # :771
        kxt = (HTE - N(HTE)) * tr  # :773
        w2 = 0.5 * (kxt + W(kxt))  # :774
        dxkx = (S(w2) - w2) * dxur  # :775
        w2 = 0.5 * (kxt + S(kxt))  # :777
        dykx = (W(w2) - w2) * dyur  # :778
        kyt = (HTN - E(HTN)) * tr  # :780
        w2 = 0.5 * (kyt + S(kyt))  # :781
        dyky = (W(w2) - w2) * dyur  # :782
        w2 = 0.5 * (kyt + W(kyt))  # :784
        dxky = (S(w2) - w2) * dxur  # :785
        dum = -(dxkx + dyky + 2 * (kxu * kxu + kyu * kyu))  # :787-789
        self.DMC = dxky - dykx  # :790-792
        self.DME = (2 * kyu) / (HTN + S(HTN))  # :795-797
        self.DMN = -(2 * kxu) / (HTE + W(HTE))  # :799-801
        duc = -(self.DUN + self.DUS + self.DUE + self.DUW)  # :803
        self.DMW = -self.DME  # :804
        self.DMS = -self.DMN  # :805
        self.cc = duc + dum  # :809

    def _component(self, p, q):
        # :811-822 (and :824-835 with the roles of u and v exchanged, same signs)
        return 1 * (self.cc * p + self.DUN * N(p) + self.DUS * S(p) + self.DUE * E(p) + self.DUW * W(p)
                    + self.DMC * q + self.DMN * N(q) + self.DMS * S(q) + self.DME * E(q) + self.DMW * W(q))

    def apply(self, u, v):
        u, v = z(u), z(v)  # :743-744
        return self._component(u, v), self._component(v, u)


# grid-var names in the reference's positional order (``required_grid_args``, kernels.py:58-63)
OPERATORS = {
    "REGULAR": (Regular, []),
    "REGULAR_AREA_WEIGHTED": (Regular, ["area"]),
    "REGULAR_WITH_LAND": (RegularWithLand, ["wet_mask"]),
    "REGULAR_WITH_LAND_AREA_WEIGHTED": (RegularWithLand, ["area", "wet_mask"]),
    "IRREGULAR_WITH_LAND": (IrregularWithLand, ["wet_mask", "dxw", "dyw", "dxs", "dys", "area", "kappa_w", "kappa_s"]),
    "MOM5U": (MOM5U, ["wet_mask", "dxt", "dyt", "dxu", "dyu", "area_u"]),
    "MOM5T": (MOM5T, ["wet_mask", "dxt", "dyt", "dxu", "dyu", "area_t"]),
    "TRIPOLAR_REGULAR_WITH_LAND_AREA_WEIGHTED": (TripolarRegular, ["area", "wet_mask"]),
    "TRIPOLAR_POP_WITH_LAND": (TripolarPOP, ["wet_mask", "dxe", "dye", "dxn", "dyn", "tarea"]),
    "VECTOR_C_GRID": (VectorCGrid, ["wet_mask_t", "wet_mask_q", "dxT", "dyT", "dxCu", "dyCu", "dxCv", "dyCv",
                                    "dxBu", "dyBu", "area_u", "area_v", "kappa_iso", "kappa_aniso"]),
    "VECTOR_B_GRID": (VectorBGrid, ["DXU", "DYU", "HUS", "HUW", "HTE", "HTN", "UAREA", "TAREA"]),
}
AREA_WEIGHTED = {"REGULAR_AREA_WEIGHTED", "REGULAR_WITH_LAND_AREA_WEIGHTED",
                 "TRIPOLAR_REGULAR_WITH_LAND_AREA_WEIGHTED"}


def required_grid_vars(grid_type):
    return list(OPERATORS[grid_type][1])


def make_operator(grid_type, grid_vars):
    cls, names = OPERATORS[grid_type]
    if set(names) != set(grid_vars):
        raise ValueError(f"Provided `grid_vars` {list(grid_vars)} do not match expected {names}")
    return cls(**{k: grid_vars[k] for k in names})


def laplacian(grid_type, grid_vars, *fields):
    return make_operator(grid_type, grid_vars).apply(*fields)


# ------------------------------------------------------------------ filter specification

FilterSpec = namedtuple("FilterSpec", ["n_steps", "s_max", "p", "dx_min_sq"])

_N_STEPS_PARAMS = {  # filter.py:28-37: (offset, factor, exponent) by shape and ndim
    "GAUSSIAN": {1: (0.8, 0.0, 1), 2: (1.1, 0.0, 1)},
    "TAPER": {1: (2.2, 0.6, 2.5), 2: (3.2, 0.7, 2.7)},
}


def n_steps_default(ndim, filter_shape, filter_scale, dx_min, transition_width):
    """filter.py:74-89"""
    off, fac, expo = _N_STEPS_PARAMS[filter_shape][ndim]
    return max(int(np.ceil((off + fac * (np.pi / transition_width) ** expo) * (filter_scale / dx_min))), 3)


def target_function(filter_shape, s_max, filter_scale, transition_width):
    """filter.py:47-65"""
    if filter_shape == "GAUSSIAN":
        return lambda t: np.exp(-(s_max * (t + 1) / 2) * filter_scale ** 2 / 24)
    from scipy.interpolate import PchipInterpolator

    fk = PchipInterpolator(
        np.array([0, 2 * np.pi / (transition_width * filter_scale), 2 * np.pi / filter_scale, 8 * np.sqrt(s_max)]),
        np.array([1, 1, 0, 0]))
    return lambda t: fk(np.sqrt((t + 1) * (s_max / 2)))


def filter_spec(filter_scale, dx_min, filter_shape, transition_width=np.pi, ndim=2, n_steps=0):
    """filter.py:99-151: Galerkin projection of the target onto Chebyshev polynomials."""
    n = n_steps
    M = (np.pi / 2) * (2 * np.eye(n - 1) - np.diag(np.ones(n - 3), 2) - np.diag(np.ones(n - 3), -2))  # :109-113
    M[0, 0] = 3 * np.pi / 2  # :114
    s_max = ndim * (2 / dx_min) ** 2  # :121
    F = target_function(filter_shape, s_max, filter_scale, transition_width)
    pts, wts = np.polynomial.chebyshev.chebgauss(n + 1)  # :128
    resid = F(pts) - ((1 - pts) / 2 + F(1) * (pts + 1) / 2)  # :135
    b = np.zeros(n - 1)
    for i in range(n - 1):  # :129-136
        sel = np.zeros(n + 1)
        sel[i], sel[i + 2] = 1, -1
        phi = np.polynomial.chebyshev.chebval(pts, sel)
        b[i] = np.sum(wts * phi * resid)
    c_hat = np.linalg.solve(M, b)  # :139
    p = np.zeros(n + 1)  # :141-147
    p[0] = c_hat[0] + (1 + F(1)) / 2
    p[1] = c_hat[1] - (1 - F(1)) / 2
    for i in range(2, n - 1):
        p[i] = c_hat[i] - c_hat[i - 2]
    p[n - 1] = -c_hat[n - 3]
    p[n] = -c_hat[n - 2]
    return FilterSpec(n, s_max, p, dx_min ** 2)


def resolve_n_steps(filter_scale, dx_min, filter_shape="GAUSSIAN", transition_width=np.pi, ndim=2, n_steps=0):
    """filter.py:352-369"""
    if ndim > 2:
        if n_steps < 3:
            raise ValueError("When ndim > 2, you must set n_steps manually")
        return n_steps
    return n_steps if n_steps >= 3 else n_steps_default(ndim, filter_shape, filter_scale, dx_min, transition_width)


# ------------------------------------------------------------------ Chebyshev step loop


def apply_filter(grid_type, grid_vars, fields, filter_scale, dx_min, filter_shape="GAUSSIAN",
                 transition_width=np.pi, ndim=2, n_steps=0):
    """filter.py:154-214 (scalar) and :217-291 (vector): returns the filtered field / (u, v)."""
    n = resolve_n_steps(filter_scale, dx_min, filter_shape, transition_width, ndim, n_steps)
    spec = filter_spec(filter_scale, dx_min, filter_shape, transition_width, ndim, n)
    op = make_operator(grid_type, grid_vars)
    return run_recurrence(op, spec, fields)


def run_recurrence(op, spec, fields):
    c = 2 / spec.s_max if op.is_dimensional else 2 / (spec.s_max * spec.dx_min_sq)  # :170-173
    p = spec.p
    if op.ncomp == 1:
        def A(x):
            return -x - c * op.apply(x)  # :169-173

        bar = op.prepare(fields[0].copy())  # :185-189
        t2 = bar.copy()  # :191
        t1 = A(bar)  # :192-194
        bar = p[0] * t2 + p[1] * t1  # :195
        for i in range(2, spec.n_steps + 1):  # :196-206
            t0 = 2 * A(t1) - t2
            bar += p[i] * t0
            t2, t1 = t1, t0
        return op.finalize(bar)  # :210

    def A2(x, y):
        lx, ly = op.apply(x, y)  # :233
        return -x - c * lx, -y - c * ly  # :235-236

    ubar, vbar = op.prepare(fields[0].copy(), fields[1].copy())  # :250-255
    u2, v2 = ubar.copy(), vbar.copy()  # :257-258
    u1, v1 = A2(ubar, vbar)  # :259-265
    ubar = p[0] * u2 + p[1] * u1  # :266
    vbar = p[0] * v2 + p[1] * v1  # :267
    for i in range(2, spec.n_steps + 1):  # :268-283
        u0, v0 = A2(u1, v1)
        u0 = 2 * u0 - u2
        v0 = 2 * v0 - v2
        ubar += p[i] * u0
        vbar += p[i] * v0
        u2, u1, v2, v1 = u1, u0, v1, v0
    return op.finalize(ubar, vbar)  # :287
