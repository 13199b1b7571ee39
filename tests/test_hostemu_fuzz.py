"""Randomised emulator-vs-oracle sweep (tests/tools/fuzz_hostemu.py) at a size that fits the CPU suite: grid type,
shape (odd widths, below / above the fused tile), batch, dtype, random land, NaNs, step count and the
steps-per-block cap are drawn at random; 6000 cases of the same generator were run clean while developing."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "tools"))
import fuzz_hostemu  # noqa: E402


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_emulator_matches_oracle_on_random_cases(seed):
    rng = np.random.default_rng(1000 + seed)
    failures = [m for m in (fuzz_hostemu.one_case(rng, k) for k in range(60)) if m]
    assert not failures, failures


def test_strided_fields_and_planes_match_contiguous():
    """C ABI: row pitch > nx, padded batch strides, unaligned base pointers (scalar forms of the vector kernels,
    one-step path instead of the fused one) give the results of the contiguous call and never write the padding."""
    import fuzz_pitch
    rng = np.random.default_rng(77)
    failures = []
    for k in range(80):
        try:
            msg = fuzz_pitch.one_case(rng, k)
        except Exception as exc:  # noqa: BLE001
            msg = f"#{k} raised {type(exc).__name__}: {exc}"
        if msg:
            failures.append(msg)
    assert not failures, failures


def test_c_grid_components_must_share_a_pitch():
    from gcm_filters_b200 import GridType, _cabi
    from gcm_filters_b200.kernels import ALL_KERNELS
    from hostemu_util import EmuPlan
    from oracle import fixtures
    (u, v), gv = fixtures.fixture("VECTOR_C_GRID", (12, 20))
    plan = EmuPlan(ALL_KERNELS[GridType.VECTOR_C_GRID](**gv), np.float64, 12, 20)
    wide = np.zeros((12, 24))
    wide[:, :20] = v
    out = [np.zeros((12, 20)), np.zeros((12, 20))]
    fin = [(u.ctypes.data, 20, 240), (wide.ctypes.data, 24, 288)]
    with pytest.raises(_cabi.GcmfError, match="share one row pitch"):
        plan.lib.laplacian(plan.h, 1, fin, [(o.ctypes.data, 20, 240) for o in out])


def test_neighbour_barrier_protocol_model():
    """tests/tools/sync_model.py: the mbarrier protocol of fused_kernel<FLUX> (and its EDGEREFILL variant) under
    random skewed schedules -- no stale or overwritten tile rows, no deadlock; the rejected LATEWAIT relaxation must
    be caught (so the checker is known to bite)."""
    import random
    import sync_model
    rng = random.Random(5)
    for edge in (0, 1):
        for skip in (0, 1):
            for k in (1, 2, 3, 4):
                for _ in range(12):
                    sync_model.Model(k, 3, edge, 0, skip, rng).run()
    caught = 0
    for _ in range(40):
        try:
            sync_model.Model(4, 3, 0, 1, 0, rng).run()
        except sync_model.Violation:
            caught += 1
    assert caught > 0
