"""CPU tests of the host side: Filter validation / messages (reference tests/test_filter.py),
FilterSpec known answers, operator registry (tests/test_kernels.py:39-61), grid validation errors
(tests/test_kernels.py:68-106, 189-221) and the exported C ABI."""
import ctypes
import os
import re

import numpy as np
import pytest

import gcm_filters_b200 as gf
from gcm_filters_b200 import Filter, FilterShape, GridType, required_grid_vars, _cabi
from gcm_filters_b200.filter import FilterSpec, _compute_filter_spec, _compute_n_steps_default
from gcm_filters_b200.kernels import ALL_KERNELS, AreaWeightedMixin, BaseScalarLaplacian, BaseVectorLaplacian
from oracle import fixtures, np_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_public_names():
    # the reference's four names (gcm_filters/__init__.py:11-15) + the device-list knob it has no counterpart for
    assert set(gf.__all__) == {"Filter", "FilterShape", "GridType", "required_grid_vars", "set_devices"}
    assert [g.name for g in GridType] == [
        "REGULAR", "REGULAR_AREA_WEIGHTED", "REGULAR_WITH_LAND", "REGULAR_WITH_LAND_AREA_WEIGHTED",
        "IRREGULAR_WITH_LAND", "MOM5U", "MOM5T", "TRIPOLAR_REGULAR_WITH_LAND_AREA_WEIGHTED",
        "TRIPOLAR_POP_WITH_LAND", "VECTOR_C_GRID", "VECTOR_B_GRID"]
    assert [s.name for s in FilterShape] == ["GAUSSIAN", "TAPER"]


@pytest.mark.parametrize("g", fixtures.SCALAR_GRIDS + fixtures.VECTOR_GRIDS)
def test_required_grid_vars_and_flags(g):
    # same names in the same positional order as the reference operator (oracle table is pinned to it)
    assert required_grid_vars(GridType[g]) == np_oracle.required_grid_vars(g)
    cls = ALL_KERNELS[GridType[g]]
    dimensional = {"IRREGULAR_WITH_LAND", "MOM5U", "MOM5T", "TRIPOLAR_POP_WITH_LAND", "VECTOR_C_GRID", "VECTOR_B_GRID"}
    assert cls.is_dimensional == (g in dimensional)
    assert issubclass(cls, BaseVectorLaplacian) == (g in fixtures.VECTOR_GRIDS)
    assert issubclass(cls, BaseScalarLaplacian) == (g not in fixtures.VECTOR_GRIDS)
    assert issubclass(cls, AreaWeightedMixin) == (g in np_oracle.AREA_WEIGHTED)


def test_filter_spec_kats():
    # reference tests/test_filter.py:23-85
    f = Filter(filter_scale=10.0, dx_min=1.0, filter_shape=FilterShape.GAUSSIAN, transition_width=np.pi, ndim=2,
               grid_vars={})
    assert f.filter_spec.n_steps == 11 and f.filter_spec.s_max == 8.0 and f.filter_spec.dx_min_sq == 1.0
    np.testing.assert_allclose(f.filter_spec.p, [0.09887381, -0.19152534, 0.1748326, -0.14975371, 0.12112337,
                                                 -0.09198484, 0.0662522, -0.04479323, 0.02895827, -0.0173953,
                                                 0.00995974, -0.00454758], rtol=1e-7, atol=1e-7)
    f = Filter(filter_scale=2.0, dx_min=1.0, filter_shape=FilterShape.TAPER, transition_width=np.pi, ndim=1,
               grid_vars={})
    assert f.filter_spec.n_steps == 6 and f.filter_spec.s_max == 4.0
    np.testing.assert_allclose(f.filter_spec.p, [0.83380304, -0.23622724, -0.06554041, 0.01593978, 0.00481014,
                                                 -0.00495532, 0.00168445], rtol=1e-7, atol=1e-7)
    assert isinstance(f.filter_spec, FilterSpec)


def test_filter_spec_matches_captured_reference(ref_outputs):
    keys = [k for k in ref_outputs if k.startswith("spec/")]
    assert keys
    for k in keys:
        _, shape, ndim, ratio, tw = k.split("/")
        ndim, ratio, tw = int(ndim), float(ratio), float(tw)
        tw = np.pi if abs(tw - np.pi) < 1e-3 else tw
        n = _compute_n_steps_default(ndim, FilterShape[shape], ratio * 0.9, 0.9, tw)
        s = _compute_filter_spec(ratio * 0.9, 0.9, FilterShape[shape], tw, ndim, n)
        ref = ref_outputs[k]
        assert s.n_steps == int(ref[0]) and s.s_max == ref[1] and s.dx_min_sq == ref[2]
        assert np.array_equal(s.p, ref[3:])  # bit-identical coefficients


def test_default_n_steps_larger_equal_3():
    assert _compute_n_steps_default(2, FilterShape.GAUSSIAN, 1.5, 1, np.pi) >= 3


def test_filter_argument_errors():
    # reference tests/test_filter.py:131-169
    (f,), gv = fixtures.fixture("REGULAR_WITH_LAND_AREA_WEIGHTED", (16, 24))
    with pytest.raises(ValueError, match=r"Provided Laplacian .*"):
        Filter(filter_scale=3.0, dx_min=2.0, grid_type=GridType.REGULAR_WITH_LAND_AREA_WEIGHTED, grid_vars=gv)
    with pytest.raises(ValueError, match=r"Transition width .*"):
        Filter(filter_scale=3.0, dx_min=1.0, transition_width=1, grid_type=GridType.REGULAR)
    with pytest.raises(ValueError, match=r"When ndim > 2, you .*"):
        Filter(filter_scale=3.0, dx_min=1.0, ndim=3, grid_type=GridType.REGULAR)
    with pytest.warns(UserWarning, match=r"You have set n_steps .*"):
        Filter(filter_scale=30.0, dx_min=1.0, n_steps=3, grid_type=GridType.REGULAR)
    with pytest.raises(ValueError, match=r"Provided `grid_vars` .*"):
        Filter(filter_scale=3.0, dx_min=1.0, grid_type=GridType.REGULAR_WITH_LAND, grid_vars={})
    with pytest.raises(ValueError, match=r"Provided `grid_vars` .*"):
        Filter(filter_scale=3.0, dx_min=1.0, grid_type=GridType.REGULAR, grid_vars={"wet_mask": gv["wet_mask"]})
    flt = Filter(filter_scale=3.0, dx_min=1.0, grid_type=GridType.REGULAR)
    assert "grid_vars" not in repr(flt) and "n_steps=" in repr(flt)
    with pytest.raises(ValueError, match=r".* is a scalar Laplacian.*"):
        flt.apply_to_vector(f, f, dims=["y", "x"])
    (u, v), gvv = fixtures.fixture("VECTOR_C_GRID", (16, 24))
    fv = Filter(filter_scale=3.0 * gvv["dxT"].min(), dx_min=gvv["dxT"].min(), grid_type=GridType.VECTOR_C_GRID,
                grid_vars=gvv)
    with pytest.raises(ValueError, match=r".* is a vector Laplacian.*"):
        fv.apply(u, dims=["y", "x"])
    with pytest.raises(AssertionError):
        # dims must name exactly two dimensions (filter.py:476); checked before any device work
        Filter(filter_scale=3.0, dx_min=1.0)._apply_to_dataarray(f, dims=["y"])


def test_grid_validation_errors():
    # reference tests/test_kernels.py:68-106
    (f,), gv = fixtures.fixture("IRREGULAR_WITH_LAND", (16, 24))
    cls = ALL_KERNELS[GridType.IRREGULAR_WITH_LAND]
    with pytest.raises(ValueError, match=r"There are kappa_w.*"):
        cls(**dict(gv, kappa_w=gv["kappa_w"] * 2))
    with pytest.raises(ValueError, match=r"There are kappa_s.*"):
        cls(**dict(gv, kappa_s=gv["kappa_s"] * 2))
    with pytest.raises(ValueError, match=r"At least one place*"):
        cls(**dict(gv, kappa_w=gv["kappa_w"] * 0.5, kappa_s=gv["kappa_s"] * 0.5))
    # reference tests/test_kernels.py:189-221
    for g in ("TRIPOLAR_REGULAR_WITH_LAND_AREA_WEIGHTED", "TRIPOLAR_POP_WITH_LAND"):
        (f,), gv = fixtures.fixture(g, (16, 24))
        with pytest.raises(AssertionError, match=r"Wet mask requires .*"):
            ALL_KERNELS[GridType[g]](**dict(gv, wet_mask=np.ones_like(gv["wet_mask"])))
    (f,), gv = fixtures.fixture("TRIPOLAR_POP_WITH_LAND", (16, 24))
    cls = ALL_KERNELS[GridType.TRIPOLAR_POP_WITH_LAND]
    with pytest.raises(AssertionError, match=r"Northernmost row of dxn.*"):
        cls(**dict(gv, dxn=fixtures.metric((16, 24), 11)))
    with pytest.raises(AssertionError, match=r"Northernmost row of dyn.*"):
        cls(**dict(gv, dyn=fixtures.metric((16, 24), 12)))
    cls(**gv)


def test_mask_dtypes_are_equivalent():
    # the reference breaks on bool / unsigned masks; here every dtype gives the same planes
    (f,), gv = fixtures.fixture("REGULAR_WITH_LAND", (16, 24))
    cls = ALL_KERNELS[GridType.REGULAR_WITH_LAND]
    base = cls(**gv)._planes.mask
    for dt in (bool, np.uint8, np.int32, np.float32):
        assert np.array_equal(cls(wet_mask=gv["wet_mask"].astype(dt))._planes.mask, base)


def test_c_abi_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "gcmf.h")).read()
    declared = set(re.findall(r"\b(gcmf_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_cabi.EXPORTS)
    if not os.path.isfile(_cabi.LIB_PATH):
        from gcm_filters_b200 import build
        build.build()
    lib = ctypes.CDLL(_cabi.LIB_PATH)  # loads without a GPU (no compute call is made)
    for name in declared:
        assert hasattr(lib, name), name
    wrapped = _cabi.Library(_cabi.LIB_PATH)
    assert wrapped.lib.gcmf_version() == 1 and wrapped.lib.gcmf_sm_arch() == 100


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    flt = Filter(filter_scale=4.0, dx_min=1.0)
    with pytest.raises(_cabi.GcmfError, match=r"no CPU fallback"):
        flt.apply(np.zeros((8, 8)), dims=["y", "x"])


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(_cabi.GcmfError, match=r"has not been built"):
        _cabi.Library(str(tmp_path / "libgcmf.so"))


def test_c_abi_from_plain_c(tmp_path):
    """include/gcmf.h is valid C99 and libgcmf.so links into a C program with no C++ / Python / torch in sight:
    tests/cabi/consumer.c makes version queries and argument-check calls (nothing that needs a GPU)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    if not os.path.isfile(_cabi.LIB_PATH):
        from gcm_filters_b200 import build
        build.build()
    exe = str(tmp_path / "consumer")
    libdir = os.path.dirname(_cabi.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cabi", "consumer.c"), "-o", exe, "-L", libdir, "-l:libgcmf.so",
                    f"-Wl,-rpath,{libdir}"], check=True)
    res = subprocess.run([exe], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "arch=100" in res.stdout and "unknown op 99" in res.stdout and res.stdout.strip().endswith("ok")


def test_c_driver_of_the_abi_against_the_emulator(tmp_path):
    """tests/cabi/gpu_vs_emu.c in its --emu-only mode (emulator against itself): keeps the plain-C GPU checker
    building and its nine cases running; on a GPU box the same binary compares libgcmf.so with the emulator."""
    import shutil
    import subprocess
    import sys
    cuda_inc = "/usr/local/cuda/include"
    if shutil.which("gcc") is None or not os.path.isfile(os.path.join(cuda_inc, "cuda_runtime_api.h")):
        pytest.skip("needs gcc and the CUDA runtime headers")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from hostemu_util import emu_library
    emu_library()
    exe = str(tmp_path / "gpu_vs_emu")
    subprocess.run(["gcc", "-std=c99", "-O1", "-Wall", "-Werror", os.path.join(ROOT, "tests", "cabi", "gpu_vs_emu.c"),
                    "-I", os.path.join(ROOT, "include"), "-I", cuda_inc, "-L", "/usr/local/cuda/lib64",
                    "-Wl,-rpath,/usr/local/cuda/lib64", "-lcudart", "-ldl", "-lm", "-o", exe], check=True)
    res = subprocess.run([exe, "--emu-only"], cwd=ROOT, capture_output=True, text=True)
    assert res.returncode == 0 and "all cases bit-identical" in res.stdout, res.stdout + res.stderr
