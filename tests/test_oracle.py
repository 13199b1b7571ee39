"""Pin the CPU oracle (oracle/np_oracle.py) to the reference: its 18 golden arrays, its FilterSpec
known-answer tests, outputs of the live reference captured in tests/golden, and -- when the
reference tree is present (build container only) -- the live reference itself, bit for bit."""
import os

import numpy as np
import pytest

from oracle import fixtures, np_oracle, ref_loader

ALL_GRIDS = fixtures.SCALAR_GRIDS + fixtures.VECTOR_GRIDS
GOLDEN_GRIDS = fixtures.GOLDEN_SCALAR_GRIDS + fixtures.VECTOR_GRIDS
GAUSS8 = dict(filter_scale=8.0, dx_min=1.0, n_steps=0, filter_shape="GAUSSIAN")


def _stack(x):
    return np.stack(x) if isinstance(x, tuple) else x


def vec_args(g, gv, fa):
    fa = dict(fa)
    if g in fixtures.VECTOR_GRIDS:
        kx, ky = ("dxT", "dyT") if g == "VECTOR_C_GRID" else ("DXU", "DYU")
        dxm = float(min(gv[kx].min(), gv[ky].min()))
        fa["dx_min"] = dxm
        fa["filter_scale"] = fa["filter_scale"] * dxm
    return fa


@pytest.mark.parametrize("g", GOLDEN_GRIDS)
def test_kernel_goldens(g, reference_goldens):
    # reference tests/test_kernels_validation.py:68-75: f4 cast, assert_allclose default rtol 1e-7
    fields, gv = fixtures.fixture(g)
    out = _stack(np_oracle.laplacian(g, gv, *fields)).astype("f4")
    np.testing.assert_allclose(reference_goldens[f"kernels/{g}"], out)


@pytest.mark.parametrize("g", GOLDEN_GRIDS)
def test_filter_goldens(g, reference_goldens):
    # reference tests/test_filter_validation.py:75-93
    fields, gv = fixtures.fixture(g)
    out = _stack(np_oracle.apply_filter(g, gv, fields, **GAUSS8)).astype("f4")
    np.testing.assert_allclose(reference_goldens[f"filter/{g}"], out)


@pytest.mark.parametrize("g", ALL_GRIDS)
@pytest.mark.parametrize("tag,shape", [("mid", (64, 96)), ("odd", (37, 54))])
def test_captured_reference_outputs_bit_exact(g, tag, shape, ref_outputs):
    fields, gv = fixtures.fixture(g, shape)
    lap = _stack(np_oracle.laplacian(g, gv, *fields))
    assert np.array_equal(lap, ref_outputs[f"lap/{tag}/{g}"])
    res = _stack(np_oracle.apply_filter(g, gv, fields, **vec_args(g, gv, GAUSS8)))
    assert np.array_equal(res, ref_outputs[f"filter/{tag}/gauss8/{g}"])


@pytest.mark.parametrize("g", ["REGULAR_WITH_LAND", "IRREGULAR_WITH_LAND", "TRIPOLAR_POP_WITH_LAND",
                               "TRIPOLAR_REGULAR_WITH_LAND_AREA_WEIGHTED", "MOM5T"])
def test_nan_land_batched(g, ref_outputs):
    (f,), gv = fixtures.fixture(g, (48, 64))
    fb = np.stack([f, f[::-1].copy(), f * f])
    fb[:, gv["wet_mask"] == 0] = np.nan
    res = np_oracle.apply_filter(g, gv, (fb,), **GAUSS8)
    assert np.array_equal(res, ref_outputs[f"filter/nanbatch/gauss8/{g}"], equal_nan=True)
    assert np.array_equal(np.isnan(res), np.broadcast_to(gv["wet_mask"] == 0, res.shape))


def test_filter_spec_kats():
    # reference tests/test_filter.py:23-79
    s = np_oracle.filter_spec(10.0, 1.0, "GAUSSIAN", np.pi, 2, np_oracle.resolve_n_steps(10.0, 1.0, "GAUSSIAN"))
    assert s.n_steps == 11 and s.s_max == 8.0 and s.dx_min_sq == 1.0
    np.testing.assert_allclose(s.p, [0.09887381, -0.19152534, 0.1748326, -0.14975371, 0.12112337, -0.09198484,
                                     0.0662522, -0.04479323, 0.02895827, -0.0173953, 0.00995974, -0.00454758],
                               rtol=1e-7, atol=1e-7)
    n = np_oracle.resolve_n_steps(2.0, 1.0, "TAPER", np.pi, 1)
    s = np_oracle.filter_spec(2.0, 1.0, "TAPER", np.pi, 1, n)
    assert s.n_steps == 6 and s.s_max == 4.0
    np.testing.assert_allclose(s.p, [0.83380304, -0.23622724, -0.06554041, 0.01593978, 0.00481014, -0.00495532,
                                     0.00168445], rtol=1e-7, atol=1e-7)


def test_filter_spec_sweep(ref_outputs):
    keys = [k for k in ref_outputs if k.startswith("spec/")]
    assert len(keys) == 32
    for k in keys:
        _, shape, ndim, ratio, tw = k.split("/")
        ndim, ratio, tw = int(ndim), float(ratio), float(tw)
        tw = np.pi if abs(tw - np.pi) < 1e-3 else tw
        n = np_oracle.n_steps_default(ndim, shape, ratio * 0.9, 0.9, tw)
        s = np_oracle.filter_spec(ratio * 0.9, 0.9, shape, tw, ndim, n)
        ref = ref_outputs[k]
        assert s.n_steps == int(ref[0]) and s.s_max == ref[1] and s.dx_min_sq == ref[2]
        assert np.array_equal(s.p, ref[3:])


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("g", ALL_GRIDS)
def test_live_reference_bit_exact(g):
    fields, gv = fixtures.fixture(g, (40, 56))
    rng = np.random.default_rng(7)
    fields = tuple(f + rng.standard_normal(f.shape) for f in fields)
    lap_ref = _stack(ref_loader.ref_laplacian(g, gv, *fields))
    assert np.array_equal(_stack(np_oracle.laplacian(g, gv, *fields)), lap_ref)
    fa = vec_args(g, gv, dict(filter_scale=5.0, dx_min=1.0, filter_shape="TAPER", n_steps=12))
    res_ref, flt = ref_loader.ref_filter(g, gv, fields, **fa)
    assert np.array_equal(_stack(np_oracle.apply_filter(g, gv, fields, **fa)), _stack(res_ref))


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present (GPU box)")
def test_live_reference_bit_exact_on_random_cases():
    """tests/tools/fuzz_oracle_vs_reference.py at suite size: random grid type, shape, batch, land, NaN / inf on land,
    batched grid variables, fp32 / fp64, both filter shapes; values and dtypes must agree exactly."""
    import subprocess
    import sys
    tool = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tools", "fuzz_oracle_vs_reference.py")
    res = subprocess.run([sys.executable, tool, "21", "80"], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]


def test_oracle_validation_errors():
    # reference tests/test_kernels.py:68-106, 189-221
    (f,), gv = fixtures.fixture("IRREGULAR_WITH_LAND", (16, 24))
    bad = dict(gv, kappa_w=gv["kappa_w"] * 2)
    with pytest.raises(ValueError, match=r"There are kappa_w.*"):
        np_oracle.make_operator("IRREGULAR_WITH_LAND", bad)
    bad = dict(gv, kappa_w=gv["kappa_w"] * 0.5, kappa_s=gv["kappa_s"] * 0.5)
    with pytest.raises(ValueError, match=r"At least one place.*"):
        np_oracle.make_operator("IRREGULAR_WITH_LAND", bad)
    for g in ("TRIPOLAR_REGULAR_WITH_LAND_AREA_WEIGHTED", "TRIPOLAR_POP_WITH_LAND"):
        (f,), gv = fixtures.fixture(g, (16, 24))
        bad = dict(gv, wet_mask=np.ones_like(gv["wet_mask"]))
        with pytest.raises(AssertionError, match=r"Wet mask requires.*"):
            np_oracle.make_operator(g, bad)
    (f,), gv = fixtures.fixture("TRIPOLAR_POP_WITH_LAND", (16, 24))
    bad = dict(gv, dxn=fixtures.metric((16, 24), 11))
    with pytest.raises(AssertionError, match=r"Northernmost row of dxn.*"):
        np_oracle.make_operator("TRIPOLAR_POP_WITH_LAND", bad)
