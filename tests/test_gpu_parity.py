"""GPU parity tests: the CUDA path of gcm_filters_b200 (through the C ABI of libgcmf.so) against the
CPU oracle and the reference's golden arrays.  Run on the B200 box: pytest -m gpu."""
import numpy as np
import pytest

from gcm_filters_b200 import Filter, FilterShape, GridType, _cabi
from gcm_filters_b200.kernels import ALL_KERNELS
from oracle import fixtures, np_oracle

from conftest import rel_l2

pytestmark = pytest.mark.gpu

ALL_GRIDS = fixtures.SCALAR_GRIDS + fixtures.VECTOR_GRIDS
GOLDEN_GRIDS = fixtures.GOLDEN_SCALAR_GRIDS + fixtures.VECTOR_GRIDS
BIT_EXACT = {"REGULAR", "REGULAR_AREA_WEIGHTED", "REGULAR_WITH_LAND", "REGULAR_WITH_LAND_AREA_WEIGHTED",
             "TRIPOLAR_REGULAR_WITH_LAND_AREA_WEIGHTED", "VECTOR_B_GRID"}
TOL64, TOL32 = 1e-12, 1e-5  # north_star: rel-L2 <= 1e-12 (fp64) / 1e-5 (fp32)


def tup(x):
    return x if isinstance(x, tuple) else (x,)


def vec_args(g, gv, fa):
    fa = dict(fa)
    if g in fixtures.VECTOR_GRIDS:
        kx, ky = ("dxT", "dyT") if g == "VECTOR_C_GRID" else ("DXU", "DYU")
        dxm = float(min(gv[kx].min(), gv[ky].min()))
        fa["dx_min"] = dxm
        fa["filter_scale"] = fa["filter_scale"] * dxm
    return fa


def make_filter(g, gv, **fa):
    fa = dict(fa)
    if isinstance(fa.get("filter_shape"), str):
        fa["filter_shape"] = FilterShape[fa["filter_shape"]]
    return Filter(grid_type=GridType[g], grid_vars=gv, **fa)


def run_filter(flt, fields):
    if len(fields) == 2:
        return tuple(flt.apply_to_vector(fields[0], fields[1], dims=["y", "x"]))
    return (flt.apply(fields[0], dims=["y", "x"]),)


def test_extension_is_the_cuda_build():
    lib = _cabi.get_library()
    assert lib.lib.gcmf_sm_arch() == 100


@pytest.mark.parametrize("shape", [(37, 54), (128, 256)])
@pytest.mark.parametrize("g", ALL_GRIDS)
def test_laplacian_vs_oracle(g, shape):
    fields, gv = fixtures.fixture(g, shape)
    lap = ALL_KERNELS[GridType[g]](**gv)
    ref = tup(np_oracle.laplacian(g, gv, *fields))
    got = tup(lap(*fields))
    for a, b in zip(got, ref):
        assert isinstance(a, np.ndarray) and a.dtype == np.float64
        if g in BIT_EXACT:
            assert np.array_equal(a, b)
        else:
            assert rel_l2(a, b) < 1e-14


@pytest.mark.parametrize("g", GOLDEN_GRIDS)
def test_reference_kernel_goldens(g, reference_goldens):
    # reference tests/test_kernels_validation.py:68-75 (f4 cast, rtol 1e-7)
    fields, gv = fixtures.fixture(g)
    got = np.stack(tup(ALL_KERNELS[GridType[g]](**gv)(*fields))).astype("f4")
    if got.shape[0] == 1:
        got = got[0]
    np.testing.assert_allclose(reference_goldens[f"kernels/{g}"], got, rtol=1e-7)


@pytest.mark.parametrize("g", GOLDEN_GRIDS)
def test_reference_filter_goldens(g, reference_goldens):
    # reference tests/test_filter_validation.py:75-93: Gaussian, filter_scale 8, dx_min 1, n_steps 0
    fields, gv = fixtures.fixture(g)
    flt = make_filter(g, gv, filter_scale=8.0, dx_min=1.0, n_steps=0, filter_shape="GAUSSIAN")
    got = np.stack(run_filter(flt, fields)).astype("f4")
    if got.shape[0] == 1:
        got = got[0]
    np.testing.assert_allclose(reference_goldens[f"filter/{g}"], got, rtol=1e-7)


@pytest.mark.parametrize("tag,shape", [("mid", (64, 96)), ("odd", (37, 54))])
@pytest.mark.parametrize("g", ALL_GRIDS)
def test_filter_vs_captured_reference(g, tag, shape, ref_outputs):
    """fp64 outputs of the live reference captured by tests/golden/make_golden.py (the only pin for MOM5U/T)."""
    fields, gv = fixtures.fixture(g, shape)
    fa = vec_args(g, gv, dict(filter_scale=8.0, dx_min=1.0, n_steps=0, filter_shape="GAUSSIAN"))
    got = np.stack(run_filter(make_filter(g, gv, **fa), fields))
    ref = ref_outputs[f"filter/{tag}/gauss8/{g}"]
    ref = ref if ref.ndim == 3 else ref[None]
    assert rel_l2(got, ref) < TOL64
    if g in BIT_EXACT:
        assert np.array_equal(got, ref)


@pytest.mark.parametrize("g", ALL_GRIDS)
def test_taper_filter_vs_oracle_f64(g):
    fields, gv = fixtures.fixture(g, (96, 130))
    fa = vec_args(g, gv, dict(filter_scale=6.0, dx_min=1.0, filter_shape="TAPER"))
    ref = tup(np_oracle.apply_filter(g, gv, fields, **fa))
    got = run_filter(make_filter(g, gv, **fa), fields)
    for a, b in zip(got, ref):
        assert rel_l2(a, b) < TOL64


@pytest.mark.parametrize("g", ALL_GRIDS)
def test_filter_f32(g):
    fields, gv = fixtures.fixture(g, (64, 128))
    fa = vec_args(g, gv, dict(filter_scale=8.0, dx_min=1.0, filter_shape="GAUSSIAN"))
    ref = tup(np_oracle.apply_filter(g, gv, fields, **fa))  # fp64 truth
    gv32 = {k: v.astype(np.float32) for k, v in gv.items()}
    got = run_filter(make_filter(g, gv32, **fa), tuple(f.astype(np.float32) for f in fields))
    for a, b in zip(got, ref):
        assert a.dtype == np.float32
        assert rel_l2(a, b) < TOL32


@pytest.mark.parametrize("g", ["REGULAR_WITH_LAND", "IRREGULAR_WITH_LAND", "TRIPOLAR_POP_WITH_LAND",
                               "TRIPOLAR_REGULAR_WITH_LAND_AREA_WEIGHTED", "MOM5T"])
def test_nan_land_batched(g, ref_outputs):
    """NaN on land in -> NaN on exactly the land cells out; wet values unaffected."""
    (f,), gv = fixtures.fixture(g, (48, 64))
    fb = np.stack([f, f[::-1].copy(), f * f])
    fb[:, gv["wet_mask"] == 0] = np.nan
    flt = make_filter(g, gv, filter_scale=8.0, dx_min=1.0)
    got = flt.apply(fb, dims=["y", "x"])
    ref = ref_outputs[f"filter/nanbatch/gauss8/{g}"]
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    assert rel_l2(got, ref) < TOL64
    if g == "MOM5T":
        return  # the reference masks MOM5 fluxes with the "wrong" neighbour (kernels.py:405-416): land leaks
    # land values must not influence wet results
    fz = np.where(np.isnan(fb), 123.0, fb)
    got2 = flt.apply(fz, dims=["y", "x"])
    wet = ~np.isnan(ref)
    assert np.array_equal(got2[wet], got[wet])


def test_batched_planes_and_4d_field():
    rng = np.random.default_rng(3)
    shape = (40, 72)
    mask = (rng.random((3,) + shape) > 0.3).astype(np.float64)
    f = rng.random((2, 3) + shape)
    flt = make_filter("REGULAR_WITH_LAND", {"wet_mask": mask}, filter_scale=4.0, dx_min=1.0)
    ref = np_oracle.apply_filter("REGULAR_WITH_LAND", {"wet_mask": mask}, (f,), filter_scale=4.0, dx_min=1.0)
    got = flt.apply(f, dims=["y", "x"])
    assert got.shape == f.shape and np.array_equal(got, ref)


def test_torch_inputs_stay_on_device():
    import torch
    (f,), gv = fixtures.fixture("IRREGULAR_WITH_LAND", (64, 96))
    flt = make_filter("IRREGULAR_WITH_LAND", gv, filter_scale=8.0, dx_min=1.0)
    ref = flt.apply(f, dims=["y", "x"])
    t = torch.as_tensor(f).cuda()
    keep = t.clone()
    out = flt.apply(t, dims=["y", "x"])
    assert out.is_cuda and torch.equal(t, keep)
    assert np.array_equal(out.cpu().numpy(), ref)
    pinned = torch.as_tensor(f).pin_memory()
    out = flt.apply(pinned, dims=["y", "x"])
    assert (not out.is_cuda) and out.is_pinned() and np.array_equal(out.numpy(), ref)


# ---- properties that hold at any size (run at benchmark-like sizes) ------------------------------
@pytest.mark.parametrize("g", ["REGULAR_WITH_LAND", "IRREGULAR_WITH_LAND", "TRIPOLAR_POP_WITH_LAND"])
def test_conservation_linearity_large(g):
    shape = (600, 900)
    (f,), gv = fixtures.fixture(g, shape)
    area = gv.get("area", gv.get("tarea", np.ones(shape)))
    lap = ALL_KERNELS[GridType[g]](**gv)
    d = lap(f)
    # reference tests/test_kernels.py:15-36: sum(area * Lap f) == 0
    assert abs(np.sum(area * d)) < 1e-12 * shape[0] * shape[1]
    flt = make_filter(g, gv, filter_scale=12.0, dx_min=1.0)
    rng = np.random.default_rng(11)
    h = rng.standard_normal(shape)
    a, b, ab = flt.apply(f, None), flt.apply(h, None), flt.apply(2.0 * f - 3.0 * h, None)
    wet = gv["wet_mask"] == 1
    assert rel_l2(ab[wet], (2.0 * a - 3.0 * b)[wet]) < 1e-12
    # reference tests/test_filter.py:118-125: integral conserved, variance reduced
    np.testing.assert_allclose(np.sum((a * area)[wet]), np.sum((f * area)[wet]), rtol=1e-9)
    assert np.var(a[wet]) < np.var(f[wet])


def test_flux_isotropy_index_kat():
    """reference tests/test_kernels.py:109-183: a delta next to outlier spacings detects a wrong shift sign."""
    for g, names in (("IRREGULAR_WITH_LAND", ("dxw", "dyw", "dxs", "dys")),
                     ("TRIPOLAR_POP_WITH_LAND", ("dxe", "dye", "dxn", "dyn"))):
        (f,), gv = fixtures.fixture(g)
        ny, nx = f.shape
        delta = np.zeros_like(f)
        delta[99, 225] = 1.0
        for k in names + (("area",) if g == "IRREGULAR_WITH_LAND" else ("tarea",)):
            gv[k] = np.ones_like(f)
        rng = np.random.default_rng(1)
        for k in names:
            gv[k] = gv[k].copy()
            gv[k][rng.integers(0, ny, 40), rng.integers(0, nx, 40)] = 7.0
        got = ALL_KERNELS[GridType[g]](**gv)(delta)
        ref = np_oracle.laplacian(g, gv, delta)
        assert rel_l2(got, ref) < 1e-14


def test_tripolar_exchange_kat():
    # reference tests/test_kernels.py:224-245
    for g in ("TRIPOLAR_REGULAR_WITH_LAND_AREA_WEIGHTED", "TRIPOLAR_POP_WITH_LAND"):
        f, gv = fixtures.tripolar_unit_fixture(g)
        nx = f.shape[1]
        delta = np.zeros_like(f)
        delta[-1, 10] = 1
        d = ALL_KERNELS[GridType[g]](**gv)(delta)
        assert d[-2, 10] == d[-1, nx - 10 - 1]


def test_vector_solid_body_rotation():
    # reference tests/test_kernels.py:251-270: Lap(u = cos(lat), v = 0) == 0
    for g in fixtures.VECTOR_GRIDS:
        _, gv = fixtures.fixture(g)
        _, geolatCu, _, _ = fixtures.spherical_geometry()
        u = np.cos(geolatCu / 360 * 2 * np.pi)
        v = np.zeros_like(u)
        if g == "VECTOR_C_GRID":
            gv = dict(gv, wet_mask_t=np.ones_like(u), wet_mask_q=np.ones_like(u))
        du, dv = ALL_KERNELS[GridType[g]](**gv)(u, v)
        np.testing.assert_allclose(du[1:-1, :], 0.0, atol=1e-12)
        np.testing.assert_allclose(dv[1:-1, :], 0.0, atol=1e-12)


@pytest.mark.parametrize("dtype,shape", [(np.float64, (150, 380)), (np.float32, (96, 520)), (np.float64, (36, 128))])
@pytest.mark.parametrize("g", ["IRREGULAR_WITH_LAND", "MOM5U", "REGULAR_WITH_LAND", "REGULAR"])
def test_fused_steps_equal_one_step_kernels(g, dtype, shape):
    """The temporally blocked kernel must give bit-identical results to the one-step kernels."""
    from gcm_filters_b200 import engine
    (f,), gv = fixtures.fixture(g, shape)
    fb = np.stack([f, f * f, 1.0 - f]).astype(dtype)
    if "wet_mask" in gv:
        fb[:, gv["wet_mask"] == 0] = np.nan
    gvt = {k: v.astype(dtype) for k, v in gv.items()}
    flt = make_filter(g, gvt, filter_scale=10.0, dx_min=1.0)
    lib = _cabi.get_library()
    try:
        engine.set_steps_per_block(0)
        n0 = lib.launch_count()
        fused = flt.apply(fb, None)
        n_fused = lib.launch_count() - n0
        engine.set_steps_per_block(1)
        n0 = lib.launch_count()
        plain = flt.apply(fb, None)
        n_plain = lib.launch_count() - n0
        engine.set_steps_per_block(3)
        capped = flt.apply(fb, None)
    finally:
        engine.set_steps_per_block(0)
    assert n_plain == flt.n_steps and n_fused == -(-flt.n_steps // 4)
    assert np.array_equal(fused, plain, equal_nan=True)
    assert np.array_equal(capped, plain, equal_nan=True)
    ref = np_oracle.apply_filter(g, gv, (fb.astype(np.float64),), filter_scale=10.0, dx_min=1.0)
    assert rel_l2(fused, ref) < (TOL64 if dtype == np.float64 else TOL32)


@pytest.mark.parametrize("dtype,shape", [(np.float64, (150, 380)), (np.float32, (96, 520))])
@pytest.mark.parametrize("g", ["TRIPOLAR_POP_WITH_LAND", "TRIPOLAR_REGULAR_WITH_LAND_AREA_WEIGHTED"])
def test_fused_tripolar_fold(g, dtype, shape):
    """Fused tiles across the tripolar fold evolve mirrored halo rows: equal to the one-step kernels up to the
    order in which mirrored cells sum their fluxes."""
    from gcm_filters_b200 import engine
    (f,), gv = fixtures.fixture(g, shape)
    fb = np.stack([f, f * f, 1.0 - f]).astype(dtype)
    fb[:, gv["wet_mask"] == 0] = np.nan
    gvt = {k: v.astype(dtype) for k, v in gv.items()}
    flt = make_filter(g, gvt, filter_scale=10.0, dx_min=1.0)
    lib = _cabi.get_library()
    try:
        engine.set_steps_per_block(0)
        n0 = lib.launch_count()
        fused = flt.apply(fb, None)
        n_fused = lib.launch_count() - n0
        engine.set_steps_per_block(1)
        plain = flt.apply(fb, None)
    finally:
        engine.set_steps_per_block(0)
    assert n_fused == -(-flt.n_steps // 4) + (1 if "AREA" in g else 0)
    assert np.array_equal(np.isnan(fused), np.isnan(plain))
    assert rel_l2(fused, plain) < (1e-14 if dtype == np.float64 else 1e-6)
    ref = np_oracle.apply_filter(g, gv, (fb.astype(np.float64),), filter_scale=10.0, dx_min=1.0)
    assert rel_l2(fused, ref) < (TOL64 if dtype == np.float64 else TOL32)
    assert rel_l2(fused[..., -8:, :], ref[..., -8:, :]) < (TOL64 if dtype == np.float64 else TOL32)


def test_c_abi_error_reporting():
    lib = _cabi.get_library()
    with pytest.raises(_cabi.GcmfError, match="unknown op"):
        lib.plan_create(99, _cabi.GCMF_F64, 8, 8, 0, 0)
    h = lib.plan_create(_cabi.OP_FLUX, _cabi.GCMF_F64, 8, 8, _cabi.FLAG_WRAP_Y, 0)
    with pytest.raises(_cabi.GcmfError, match="out of range"):
        lib.plan_set_plane(h, 7, 1234, 8, 64, 1)
    import torch
    x = torch.zeros((1, 8, 8), dtype=torch.float64, device="cuda")
    y = torch.zeros_like(x)
    spec = [(x.data_ptr(), 8, 64)]
    with pytest.raises(_cabi.GcmfError, match="has not been set"):
        lib.laplacian(h, 1, spec, [(y.data_ptr(), 8, 64)])
    lib.plan_destroy(h)


def test_pipelined_host_path_matches_monolithic():
    """Host inputs with a long batch axis are streamed in chunks (H2D / filter / D2H overlapped)."""
    import torch
    from gcm_filters_b200 import engine
    (f,), gv = fixtures.fixture("IRREGULAR_WITH_LAND", (64, 160))
    rng = np.random.default_rng(2)
    fb = f[None] * (1 + 0.1 * rng.standard_normal((37, 1, 1)))
    fb[:, gv["wet_mask"] == 0] = np.nan
    flt = make_filter("IRREGULAR_WITH_LAND", gv, filter_scale=8.0, dx_min=1.0)
    mono = flt.apply(torch.as_tensor(fb).cuda(), None).cpu().numpy()  # device-resident: never pipelined
    assert engine._pipeline_chunk(37, 64 * 160 * 8) == 4
    piped = flt.apply(fb, None)  # numpy in, numpy out: chunks of 1, 2, 4 x 7, 3, 2, 1 slices
    assert isinstance(piped, np.ndarray) and np.array_equal(piped, mono, equal_nan=True)
    pin_in = torch.as_tensor(fb).pin_memory()
    pin_out = torch.empty_like(pin_in).pin_memory()
    res = flt.apply(pin_in, None, out=pin_out)
    assert res is pin_out and np.array_equal(pin_out.numpy(), mono, equal_nan=True)
    ref = np_oracle.apply_filter("IRREGULAR_WITH_LAND", gv, (fb,), filter_scale=8.0, dx_min=1.0)
    assert rel_l2(piped, ref) < TOL64


@pytest.mark.parametrize("g", ["IRREGULAR_WITH_LAND", "TRIPOLAR_POP_WITH_LAND"])
def test_fused_neighbour_sync_is_deterministic(g):
    """Both fused FLUX kernels order their shared-memory traffic with mbarriers instead of CTA barriers (the tile form:
    per-warp barriers between neighbouring warps -- it runs the tripolar grid here; the row-streaming form: TMA ring
    full / empty barriers + one named barrier per level and row -- it runs the periodic grid), which compute-sanitizer
    racecheck does not model.  A missing dependency would show up as run-to-run differences: 25 repetitions must agree
    bit for bit with each other and with the barrier-free one-step kernels (to rounding across the tripolar fold)."""
    import torch
    from gcm_filters_b200 import engine
    (f,), gv = fixtures.fixture(g, (480, 720))
    rng = np.random.default_rng(9)
    fb = f[None] * (1 + 0.2 * rng.standard_normal((6, 1, 1)))
    fb[:, gv["wet_mask"] == 0] = np.nan
    flt = make_filter(g, gv, filter_scale=16.0, dx_min=1.0)
    t = torch.as_tensor(fb).cuda()
    try:
        engine.set_steps_per_block(1)
        plain = flt.apply(t, None).clone()
    finally:
        engine.set_steps_per_block(0)
    first = None
    for rep in range(25):
        out = flt.apply(t, None)
        if g == "IRREGULAR_WITH_LAND":
            assert torch.equal(torch.nan_to_num(out, nan=-1.0), torch.nan_to_num(plain, nan=-1.0)), rep
        elif first is None:
            first = out.clone()
            assert rel_l2(first.cpu().numpy(), plain.cpu().numpy()) < 1e-14
        else:
            assert torch.equal(torch.nan_to_num(out, nan=-1.0), torch.nan_to_num(first, nan=-1.0)), rep


def test_full_size_pop_slice_vs_oracle_and_properties():
    """BASELINE config 3 at its full horizontal size (2400 x 3600, fp64, NaN on land), two levels: the fused
    CUDA path against the oracle (12 forced steps keep the numpy run short), plus size-independent properties."""
    cfg = fixtures.cfg3(nb=2)
    (f,), gv = cfg["fields"], cfg["grid_vars"]
    fa = dict(filter_scale=36.0, dx_min=0.9, filter_shape=FilterShape.GAUSSIAN, n_steps=12)
    with pytest.warns(UserWarning, match="n_steps below the default"):
        flt = Filter(grid_type=GridType.IRREGULAR_WITH_LAND, grid_vars=gv, **fa)
    got = flt.apply(f, None)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = np_oracle.apply_filter("IRREGULAR_WITH_LAND", gv, (f[:1],), filter_scale=36.0, dx_min=0.9, n_steps=12)
    assert np.array_equal(np.isnan(got[:1]), np.isnan(ref))
    assert rel_l2(got[:1], ref) < TOL64
    wet = gv["wet_mask"] == 1
    # the filter conserves the area integral (reference tests/test_filter.py:118-121) ...
    area = gv["area"]
    np.testing.assert_allclose(np.sum((got[1] * area)[wet]), np.sum((f[1] * area)[wet]), rtol=1e-10)
    # ... leaves a constant untouched and is linear
    const = np.where(wet, 3.25, np.nan)[None]
    np.testing.assert_allclose(flt.apply(const, None)[0][wet], 3.25, rtol=1e-13)
    mix = flt.apply(2.0 * f[:1] - 0.5 * f[1:], None)
    assert rel_l2(mix[0][wet], (2.0 * got[0] - 0.5 * got[1])[wet]) < 1e-12


@pytest.mark.parametrize("g,dtype", [("IRREGULAR_WITH_LAND", np.float64), ("REGULAR_WITH_LAND", np.float32),
                                     ("TRIPOLAR_POP_WITH_LAND", np.float64)])
@pytest.mark.parametrize("levels", [2, 5])
def test_fused_level_slabs(g, dtype, levels, monkeypatch):
    """A CTA that loops over several levels (TMA refill of the landing tiles during the steps, mbarrier parity,
    growing progress counters): forced here because small grids get one level per CTA on their own; the benchmark
    configs run with ~30 levels per CTA."""
    from gcm_filters_b200 import engine
    shape = (96, 520)
    (f,), gv = fixtures.fixture(g, shape)
    rng = np.random.default_rng(levels)
    fb = (f[None] * (1 + 0.3 * rng.standard_normal((11, 1, 1)))).astype(dtype)
    fb[:, gv["wet_mask"] == 0] = np.nan
    gvt = {k: v.astype(dtype) for k, v in gv.items()}
    flt = make_filter(g, gvt, filter_scale=10.0, dx_min=1.0)
    import torch
    dev = torch.as_tensor(fb).cuda()  # device-resident: one launch over all 11 levels (host inputs are chunked)
    monkeypatch.setenv("GCMF_FUSED_LEVELS_PER_CTA", str(levels))
    fused = flt.apply(dev, None).cpu().numpy()
    monkeypatch.delenv("GCMF_FUSED_LEVELS_PER_CTA")
    try:
        engine.set_steps_per_block(1)
        plain = flt.apply(dev, None).cpu().numpy()
    finally:
        engine.set_steps_per_block(0)
    if g.startswith("TRIPOLAR"):
        assert np.array_equal(np.isnan(fused), np.isnan(plain)) and rel_l2(fused, plain) < 1e-14
    else:
        assert np.array_equal(fused, plain, equal_nan=True)


def test_concurrent_calls_from_host_threads():
    """dask's threaded scheduler calls filter_func from several host threads (one block each): the launch
    sequences share the device workspace and must not interleave."""
    from concurrent.futures import ThreadPoolExecutor
    (f,), gv = fixtures.fixture("IRREGULAR_WITH_LAND", (96, 264))
    flt = make_filter("IRREGULAR_WITH_LAND", gv, filter_scale=10.0, dx_min=1.0)
    rng = np.random.default_rng(12)
    blocks = [f[None] * (1 + 0.2 * rng.standard_normal((3, 1, 1))) for _ in range(8)]
    serial = [flt.apply(b, None) for b in blocks]
    with ThreadPoolExecutor(max_workers=4) as pool:
        threaded = list(pool.map(lambda b: flt.apply(b, None), blocks * 3))
    for k, out in enumerate(threaded):
        assert np.array_equal(out, serial[k % len(blocks)], equal_nan=True)


def test_operator_protocol_prepare_call_finalize():
    """prepare / __call__ / finalize of the area-weighted operators (reference kernels.py:89-104) run on the device."""
    for g in ("REGULAR_AREA_WEIGHTED", "REGULAR_WITH_LAND_AREA_WEIGHTED", "TRIPOLAR_REGULAR_WITH_LAND_AREA_WEIGHTED"):
        (f,), gv = fixtures.fixture(g, (48, 72))
        lap = ALL_KERNELS[GridType[g]](**gv)
        lib = _cabi.get_library()
        n0 = lib.launch_count()
        x = lap.prepare(f)
        assert lib.launch_count() == n0 + 1 and np.array_equal(x, f * gv["area"])
        assert np.array_equal(lap.finalize(x), (f * gv["area"]) / gv["area"])
        op = np_oracle.make_operator(g, gv)
        assert np.array_equal(lap(x), op.apply(op.prepare(f)))
    (f,), gv = fixtures.fixture("REGULAR_WITH_LAND", (48, 72))
    lap = ALL_KERNELS[GridType.REGULAR_WITH_LAND](**gv)
    assert lap.prepare(f) is f and lap.finalize(f) is f  # identity for the other operators (kernels.py:47-54)


@pytest.mark.parametrize("g,shape", [("VECTOR_C_GRID", (96, 488)), ("VECTOR_C_GRID", (45, 100)), ("VECTOR_C_GRID", (40, 24)),
                                     ("VECTOR_B_GRID", (64, 128)), ("IRREGULAR_WITH_LAND", (96, 264))])
def test_peer_banded_single_rank_is_its_own_neighbour(g, shape):
    """PeerBandedFilter with one rank (ordinary device memory, the band wraps onto itself): the step kernels' HALO forms
    -- ghost rows pushed by the kernel, flag wait / signal, border bands scheduled first -- on one GPU, against the
    whole-grid one-step kernels, bit for bit.  488 columns take the TMA-pipelined C-grid kernel, 100 the marching one,
    24 the tiled one."""
    from gcm_filters_b200 import engine
    from gcm_filters_b200.scheduler import PeerBandedFilter
    fields, gv = fixtures.fixture(g, shape)
    fields = tuple(np.stack([f, 1.0 - f * f]) for f in fields)
    fa = vec_args(g, gv, dict(filter_scale=6.0, dx_min=1.0))
    flt = make_filter(g, gv, **fa)
    try:
        engine.set_steps_per_block(1)
        single = run_filter(flt, fields)
    finally:
        engine.set_steps_per_block(0)
    pbf = PeerBandedFilter(flt, 0, 1)
    for _ in range(2):  # a second epoch on the reused buffers
        outs, (j0, j1) = pbf.apply(*fields)
    pbf.close()
    assert (j0, j1) == (0, shape[0])
    for o, s in zip(outs, single):
        assert np.array_equal(o, s, equal_nan=True)


@pytest.mark.parametrize("kernel", ["tiled", "march", "tma"])
def test_cgrid_kernel_forms_are_bit_identical(kernel):
    """The three device forms of the C-grid operator (tiled stress tiles, register marching, TMA-pipelined rows) run
    the same expressions in the same order: identical bits, Laplacian and filter, odd and aligned widths."""
    import subprocess
    import sys
    code = (
        "import numpy as np, sys\n"
        "from gcm_filters_b200 import Filter, GridType\n"
        "from gcm_filters_b200.kernels import ALL_KERNELS\n"
        "from oracle import fixtures\n"
        "out = {}\n"
        "for shape in ((70, 488), (37, 250), (33, 54)):\n"
        "    (u, v), gv = fixtures.fixture('VECTOR_C_GRID', shape)\n"
        "    lu, lv = ALL_KERNELS[GridType.VECTOR_C_GRID](**gv)(u, v)\n"
        "    dxm = float(min(gv['dxT'].min(), gv['dyT'].min()))\n"
        "    flt = Filter(filter_scale=5 * dxm, dx_min=dxm, grid_type=GridType.VECTOR_C_GRID, grid_vars=gv)\n"
        "    fu, fv = flt.apply_to_vector(np.stack([u, v * u]), np.stack([v, u - v]), None)\n"
        "    out[str(shape)] = np.stack([lu, lv]); out['f' + str(shape)] = np.stack([fu, fv])\n"
        "np.savez(sys.argv[1], **out)\n")
    import os
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        res = {}
        for k in ("tiled", kernel):
            path = os.path.join(tmp, k + ".npz")
            env = dict(os.environ, GCMF_CGRID_KERNEL=k, PYTHONPATH=os.pathsep.join(sys.path))
            subprocess.run([sys.executable, "-c", code, path], check=True, env=env, timeout=300)
            res[k] = dict(np.load(path))
        for key in res["tiled"]:
            assert np.array_equal(res["tiled"][key], res[kernel][key], equal_nan=True), (kernel, key)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape,nb,n_steps", [((70, 488), 2, 6), ((37, 250), 1, 7), ((64, 1000), 3, 5), ((12, 232), 1, 3),
                                               ((130, 456), 1, 9)])
@pytest.mark.parametrize("g", ["VECTOR_C_GRID", "VECTOR_B_GRID"])
def test_vector_two_step_kernel_is_bit_identical_to_one_step(g, shape, nb, n_steps, dtype, monkeypatch):
    """Temporal blocking of the vector operators (vec2_kernel: steps i and i+1 in one launch, step i+1 marching one row
    behind step i): same expressions in the same order as two one-step launches -> identical bits.  Even and odd step
    counts (a trailing one-step LAST launch), one and several column strips, bands shorter than the priming depth,
    NaNs in the input, both dtypes; half as many launches as steps."""
    from gcm_filters_b200 import engine
    (u, v), gv = fixtures.fixture(g, shape)
    rng = np.random.default_rng(5)
    us = np.stack([u * (1 + 0.1 * k) + 0.01 * rng.standard_normal(shape) for k in range(nb)]).astype(dtype)
    vs = np.stack([v - 0.2 * k * u for k in range(nb)]).astype(dtype)
    us[0, 3:5, 7:9] = np.nan  # nan_to_num inside both steps
    gvt = {k: np.asarray(a).astype(dtype) for k, a in gv.items()}
    fa = vec_args(g, gv, dict(filter_scale=6.0, dx_min=1.0, n_steps=n_steps))
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        flt = make_filter(g, gvt, **fa)
    assert flt.n_steps == n_steps
    lib = _cabi.get_library()
    av = 16 // np.dtype(dtype).itemsize  # rows must split into 16-byte vectors and span one strip (28 x 8 + halo)
    blocked = shape[1] % av == 0 and (shape[0] * shape[1]) % av == 0 and shape[1] >= 224 + 2 * max(2, av)
    for rows in (None, "5", "16"):  # default band height, bands of 5 rows (shorter than the 4 priming rows + 2), 16
        if rows:
            monkeypatch.setenv("GCMF_CGRID_ROWS", rows)
        n0 = lib.launch_count()
        fu, fv = flt.apply_to_vector(us, vs, None)
        launches = lib.launch_count() - n0
        monkeypatch.delenv("GCMF_CGRID_ROWS", raising=False)
        assert launches == ((n_steps + 1) // 2 if blocked else n_steps), (launches, n_steps, blocked)
        try:
            engine.set_steps_per_block(1)
            pu, pv = flt.apply_to_vector(us, vs, None)
        finally:
            engine.set_steps_per_block(0)
        assert fu.dtype == dtype
        assert np.array_equal(fu, pu, equal_nan=True) and np.array_equal(fv, pv, equal_nan=True), rows


@pytest.mark.parametrize("g,shape", [("VECTOR_C_GRID", (96, 488)), ("VECTOR_B_GRID", (64, 232)), ("IRREGULAR_WITH_LAND", (96, 264))])
def test_fused_banded_single_rank_is_its_own_neighbour(g, shape):
    """FusedBandedFilter with one rank (the band wraps onto itself: ghost rows refreshed by local copies once per block):
    the temporally blocked kernels on a BAND plan -- physical ghost rows instead of the periodic wrap, 2 per side for
    the two-step vector kernel, 4 for the scalar tiles -- on one GPU, against the whole-grid one-step kernels, bit for
    bit.  Odd and even step counts (the vector path ends an odd count with a one-step LAST launch)."""
    from gcm_filters_b200 import engine
    from gcm_filters_b200.scheduler import FusedBandedFilter
    fields, gv = fixtures.fixture(g, shape)
    fields = tuple(np.stack([f, 1.0 - f * f]) for f in fields)
    for scale in (6.0, 7.0):
        fa = vec_args(g, gv, dict(filter_scale=scale, dx_min=1.0))
        flt = make_filter(g, gv, **fa)
        try:
            engine.set_steps_per_block(1)
            single = run_filter(flt, fields)
        finally:
            engine.set_steps_per_block(0)
        for exch in (("nccl", "push") if len(fields) == 2 else ("nccl",)):
            # "push": the exchange fused into the two-step kernel (peer stores + flags); with one rank the band pushes
            # into its own ghost rows -- the same kernels and protocol as on several GPUs
            fbf = FusedBandedFilter(flt, 0, 1, exchange=exch)
            st = fbf.stage(*fields)
            for _ in range(3):  # several runs on one staging: the flags keep growing
                bar = fbf.run(st)
            outs = tuple(bar[k].cpu().numpy() for k in range(len(fields)))
            fbf.close()
            for o, s in zip(outs, single):
                assert np.array_equal(o, s, equal_nan=True), (g, scale, flt.n_steps, exch)


def test_fused_flux_march_form_is_bit_identical_to_tile_form():
    """The row-streaming form of the fused FLUX steps (march_kernel, opt-in with GCMF_FUSED_FORM=march): same per-point
    arithmetic as the tile form and the one-step kernels -> identical bits.  Whole grids with several strips and row
    bands, level groups that do not divide the batch, step counts that end with a short block, NaN on land."""
    import os
    import subprocess
    import sys
    import tempfile
    code = (
        "import numpy as np, sys\n"
        "from gcm_filters_b200 import Filter, GridType\n"
        "from oracle import fixtures\n"
        "out = {}\n"
        "for shape, nb, scale in (((96, 264), 4, 8.0), ((150, 520), 7, 6.0), ((64, 136), 1, 9.0)):\n"
        "    (f,), gv = fixtures.fixture('IRREGULAR_WITH_LAND', shape)\n"
        "    rng = np.random.default_rng(3)\n"
        "    fb = f[None] * (1 + 0.2 * rng.standard_normal((nb, 1, 1)))\n"
        "    fb[:, gv['wet_mask'] == 0] = np.nan\n"
        "    flt = Filter(filter_scale=scale, dx_min=1.0, grid_type=GridType.IRREGULAR_WITH_LAND, grid_vars=gv)\n"
        "    out[str(shape)] = flt.apply(fb, None)\n"
        "np.savez(sys.argv[1], **out)\n")
    with tempfile.TemporaryDirectory() as tmp:
        res = {}
        for form in ("tile", "march"):
            path = os.path.join(tmp, form + ".npz")
            env = dict(os.environ, GCMF_FUSED_FORM=form, PYTHONPATH=os.pathsep.join(sys.path))
            subprocess.run([sys.executable, "-c", code, path], check=True, env=env, timeout=300)
            res[form] = dict(np.load(path))
        for key in res["tile"]:
            assert np.array_equal(res["tile"][key], res["march"][key], equal_nan=True), key
