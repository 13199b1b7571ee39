"""Host-side model of the iteration schedule of march_kernel (gcm_filters_b200/csrc/gcmf_march.cuh).

The kernel runs eight iterations per trip and compiles the trips that lie inside a band WITHOUT the CTA-uniform range
checks (template parameter STEADY).  Dropping a check that can fail is a deadlock (an mbarrier wait for a row that is
never staged) or an out-of-range store, and only shows up on the GPU for particular (band height, block length)
pairs -- e.g. a one-step block at the end of the recurrence on a band of 24 rows.  This test restates the trip
condition and every guard of `MarchConsumer::iteration` and checks, for all block lengths and a range of band
geometries, that STEADY implies each guard it removes, and that the generic iterations of a band wait for exactly
the rows the producer stages."""
import itertools

import pytest


def steady_trip(t, j0, j1, K):  # the condition in march_kernel's trip loop
    return t >= j0 + 2 * (K - 1) and t + 7 <= j1 - 1 and t + 7 <= j1 + K - 3


def test_model_matches_the_kernel_source():
    """The trip condition modelled here is the one compiled into the kernel (a textual pin: change both or neither)."""
    import os
    src = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gcm_filters_b200", "csrc", "gcmf_march.cuh")
    with open(src) as fh:
        text = fh.read()
    assert "if (t >= j0 + 2 * (K - 1) && t + 7 <= j1 - 1 && t + 7 <= j1 + K - 3) {" in text
    assert "constexpr int TI0 = 1 + PH;" in text and "const int ci = ti - 2 * K + 1;" in text


@pytest.mark.parametrize("K", [1, 2, 3, 4])
def test_steady_trips_satisfy_every_dropped_guard(K):
    for j0, ry in itertools.product((0, 24, 100), range(1, 90)):
        j1 = j0 + ry
        R0, nrows = j0 - K, ry + 2 * K           # staged rows R0 .. j1 + K - 1
        last_idx = nrows - 1
        t0, t1 = j0 - (K - 1), j1 - 1 + 2 * (K - 1)
        waited, released = set(), set()
        t = t0
        while t <= t1:
            steady = steady_trip(t, j0, j1, K)
            for ph in range(8):
                tt = t + ph
                ti = tt - R0
                guards = {
                    "row staged two ahead": ti + 2 <= last_idx,
                    "row staged": ti <= last_idx,
                    "bar row owned": j0 <= tt < j1,
                    "output row owned": j0 <= tt - 2 * (K - 1) < j1,
                    "coefficient release": 0 <= ti - 2 * K + 1 <= last_idx,
                }
                for s in range(1, K + 1):
                    r = tt - 2 * (s - 1)
                    guards[f"step {s} active"] = j0 - (K - s) <= r <= j1 - 1 + (K - s)
                    if guards[f"step {s} active"]:  # the coefficient rows a step reads exist
                        assert 0 <= ti - 2 * (s - 1) - 1 and ti - 2 * (s - 1) <= last_idx, (K, ry, tt, s)
                if steady:
                    assert all(guards.values()), (K, j0, ry, tt, [k for k, v in guards.items() if not v])
                if steady or guards["row staged two ahead"]:
                    waited.add(ti + 2)
                if steady or guards["row staged"]:
                    released.add(ti)
            t += 8
        # with the prologue's rows 0, 1, 2 every staged row is waited for exactly once it exists, none beyond
        assert waited | {0, 1, 2} == set(range(nrows)) | {0, 1, 2}, (K, j0, ry)
        assert max(waited | {0}) <= max(last_idx, 2)
        assert released <= set(range(nrows))
