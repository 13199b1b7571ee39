"""Host-side logic of the engine / Filter front-end that needs no GPU: chunk schedules of the H2D / filter / D2H
pipeline, the multi-threaded staging copy, the grid-variable fingerprint of the Laplacian cache, batch slabs."""
import numpy as np
import pytest

from gcm_filters_b200 import engine
from gcm_filters_b200.filter import _fingerprint, _LaplacianCache
from gcm_filters_b200.scheduler import batch_slabs


@pytest.mark.parametrize("nb,chunk", [(1, 1), (7, 1), (8, 1), (62, 7), (365, 45), (100, 100), (13, 4)])
def test_chunk_schedule_covers_the_batch_once(nb, chunk):
    sizes = engine._chunk_schedule(nb, chunk)
    assert sum(sizes) == nb and all(0 < s <= chunk for s in sizes)
    assert sizes == sizes[::-1] or nb % chunk  # ramp up at the front, the same ramp down at the back


def test_pipeline_chunk_is_bounded_by_bytes_and_target():
    assert engine.PIPELINE_TARGET_CHUNKS == 8 and engine.PIPELINE_LEVEL_GROUP == 2
    assert engine._pipeline_chunk(62, 2400 * 3600 * 8) == 6   # 62 // 8 = 7 -> a whole number of level groups
    assert engine._pipeline_chunk(365, 720 * 1440 * 4) == 44
    assert engine._pipeline_chunk(37, 64 * 160 * 8) == 4 and engine._pipeline_chunk(16, 1000) == 2
    assert engine._pipeline_chunk(11, 1000) == 2              # one group per chunk still leaves four chunks
    assert engine._pipeline_chunk(7, 1000) == 1               # too short for four whole groups
    assert engine._pipeline_chunk(400, 200 << 20) == 4        # the byte cap (1 GiB) wins: 5 slices fit
    assert engine._pipeline_chunk(4, 1 << 40) == 1  # a slice larger than the byte cap still moves one at a time


def test_parallel_copy_matches_numpy_and_casts():
    import torch

    rng = np.random.default_rng(3)
    src = rng.standard_normal((5, 300, 701))  # > 1 Mi elements: split over the copy threads
    dst = torch.empty((5, 300, 701), dtype=torch.float64)
    engine._parallel_copy(dst, torch.from_numpy(src))
    assert np.array_equal(dst.numpy(), src)
    dst32 = torch.empty((5, 300, 701), dtype=torch.float32)
    engine._parallel_copy(dst32, src)  # numpy source, dtype conversion on the way
    assert np.array_equal(dst32.numpy(), src.astype(np.float32))
    small = torch.empty((3, 4), dtype=torch.float64)
    engine._parallel_copy(small, np.arange(12.0).reshape(3, 4))
    assert np.array_equal(small.numpy(), np.arange(12.0).reshape(3, 4))


def test_fingerprint_sees_in_place_edits():
    m = np.ones((128, 256))
    f0 = _fingerprint(m)
    m[77, 3] = 0.0
    assert _fingerprint(m) != f0  # small arrays are covered completely
    big = np.ones((2400, 3600))
    f1 = _fingerprint(big)
    big[0, 17] = 2.0  # sampled rows include the first and the last one
    assert _fingerprint(big) != f1
    assert _fingerprint(big[:, ::2]) == _fingerprint(np.ascontiguousarray(big[:, ::2]))  # strided views
    assert _fingerprint(np.ones(5, dtype=np.float32)) != _fingerprint(np.zeros(5, dtype=np.float32))


def test_laplacian_cache_rebuilds_after_in_place_edit():
    built = []

    class Lap:
        @staticmethod
        def required_grid_args():
            return ["wet_mask"]

        def __init__(self, wet_mask):
            built.append(wet_mask.copy())

    cache = _LaplacianCache(Lap)
    m = np.ones((16, 24))
    a = cache.get((m,))
    assert cache.get((m,)) is a and len(built) == 1          # same array, same contents: reused
    m[3, 4] = 0.0
    b = cache.get((m,))
    assert b is not a and len(built) == 2 and built[1][3, 4] == 0.0  # edited in place: rebuilt, as the reference would
    assert cache.get((m.tolist(),)) is not None                  # non-ndarray input: converted copy is kept alive
    cache.clear()
    assert cache.get((m,)) is not b


def test_batch_slabs_of_the_headline_field():
    assert [hi - lo for lo, hi in batch_slabs(62, 8)] == [8, 8, 8, 8, 8, 8, 7, 7]
    assert [hi - lo for lo, hi in batch_slabs(365, 8)] == [46, 46, 46, 46, 46, 45, 45, 45]
    assert batch_slabs(1, 4) == [(0, 1), (1, 1), (1, 1), (1, 1)]
