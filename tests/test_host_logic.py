"""Host-side logic of the Python front that needs no GPU."""
import numpy as np

from therldaisyworld_b200.env import shared_cell_offsets


def brute_force_counts(pos, N):
    """The reference's scan (daisy_world_rl.py:221-229): every cell of every world, residents.sum() > 1."""
    counts = []
    for b in range(pos.shape[0]):
        c = 0
        for xx in range(N):
            for yy in range(N):
                residents = ((pos[b] == np.array([xx, yy])).mean(-1) == 1)
                c += int(residents.sum() > 1)
        counts.append(c)
    return np.array(counts)


def test_shared_cell_offsets_match_the_reference_scan():
    rng = np.random.RandomState(0)
    for N, B, n in [(3, 5, 7), (5, 4, 12), (8, 6, 10), (2, 3, 9), (7, 2, 40), (16, 3, 4), (4, 2, 1), (6, 3, 2)]:
        pos = rng.randint(N, size=(B, n, 2))
        off = shared_cell_offsets(pos, N)
        assert off.dtype == np.int32 and off.shape == (B + 1,) and off[0] == 0
        np.testing.assert_array_equal(np.diff(off), brute_force_counts(pos, N))
    # all agents on one cell: one shared cell; all distinct: none
    pos = np.zeros((2, 6, 2), dtype=np.int64)
    pos[1, :, 1] = np.arange(6)
    np.testing.assert_array_equal(shared_cell_offsets(pos, 8), [0, 1, 1])
    assert shared_cell_offsets(np.zeros((3, 0, 2), dtype=np.int64), 8).tolist() == [0, 0, 0, 0]
