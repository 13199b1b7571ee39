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
