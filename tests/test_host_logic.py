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


def test_pinned_pool_recycles_blocks_as_soon_as_the_arrays_are_dropped():
    """env._PinnedPool hands out page-locked blocks for step()'s outputs; a block must return to the pool the moment the
    caller drops the arrays carved from it (no reference cycle that would wait for the cyclic GC), and callers that keep
    everything must fall back to ordinary memory instead of pinning without bound. (Host logic only: a malloc stand-in
    replaces dw_host_alloc.)"""
    import ctypes as C
    import gc
    from therldaisyworld_b200.env import _PinnedPool
    libc = C.CDLL(None)
    libc.malloc.restype = C.c_void_p
    allocs = []

    class FakeLib:
        def dw_host_alloc(self, n, pp):
            addr = libc.malloc(n.value)
            allocs.append(addr)
            C.cast(pp, C.POINTER(C.c_void_p))[0] = addr
            return 0

        def dw_host_free(self, p):
            return 0

    pool = _PinnedPool(FakeLib())
    gc.disable()
    try:
        def views(nbytes=4096):
            buf = pool.take(nbytes)
            if buf is None:
                return None
            ptr = C.c_void_p(C.addressof(buf))
            a = np.frombuffer(buf, dtype=np.float64, count=16, offset=0).reshape(4, 4)
            b = np.frombuffer(buf, dtype=np.uint8, count=16, offset=2048).view(np.bool_)
            return ptr, a, b
        for _ in range(100):
            out = views()
        assert len(allocs) <= 2 and pool._live == 1            # two blocks alternate
        kept = [views() for _ in range(40)]
        assert sum(k is None for k in kept) == 40 - (_PinnedPool.MAX_LIVE - 1)   # beyond the limit: fall back
        del kept, out
        assert pool._live == 0
    finally:
        gc.enable()
