"""Host-side logic of the row-banded giant world (therldaisyworld_b200/banded.py) on CPU: bands backed by the NumPy
oracle (tests/band_helpers.OracleBand), exchanges over real torch.distributed gloo (2 ranks) or in-process threads
(4 bands), must reproduce the full-torus C oracle exactly -- covers, agents, lifespans."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT

N, N_AGENTS, STEPS = 128, 24, 40


def _reference(policy, actions):
    from band_helpers import full_oracle, make_state
    from therldaisyworld_b200.banded import BandedDaisyWorld
    from band_helpers import OracleBand
    light, dark, ai, st = make_state(N, N_AGENTS, seed=5, clustered=True)
    w = BandedDaisyWorld(N, N_AGENTS, band_factory=OracleBand)      # only used as a parameter carrier here
    ref = full_oracle(w, light, dark, ai, st)
    steps, done_at, ada = ref.run(STEPS, policy, actions=actions)
    return (light, dark, ai, st), ref, done_at, ada


def _actions():
    return np.random.RandomState(3).randint(9, size=(STEPS, N_AGENTS))


def _check(world, ref, done_at, ada, covers):
    np.testing.assert_array_equal(covers[0], ref.grid[0, 1])
    np.testing.assert_array_equal(covers[1], ref.grid[0, 2])
    xy, st = world.agents()
    np.testing.assert_array_equal(xy, ref.agent_indices[0])
    np.testing.assert_array_equal(st, ref.agent_states[0, :, 0])
    d, a = world.lifespans()
    assert d == int(done_at[0])
    np.testing.assert_array_equal(a, ada[0, :, 0])


def _worker(rank, world_size, port, policy, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world_size))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from band_helpers import OracleBand, make_state
    from therldaisyworld_b200.banded import BandedDaisyWorld
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    light, dark, ai, st = make_state(N, N_AGENTS, seed=5, clustered=True)
    w = BandedDaisyWorld(N, N_AGENTS, rank=rank, world_size=world_size, band_factory=OracleBand)
    w.load_state(light, dark, ai, st)
    w.run(STEPS, policy, actions=_actions() if policy == "replay" else None, chunk=16)
    q.put((rank, w.local_covers(), w.agents(), w.lifespans(), w.step_count, w.L))
    dist.destroy_process_group()


@pytest.mark.parametrize("policy", ["greedy", "replay"])
def test_two_gloo_ranks_match_full_torus_oracle(policy):
    import torch.multiprocessing as mp
    acts = _actions() if policy == "replay" else None
    _, ref, done_at, ada = _reference(policy, None if acts is None else acts[:, None, :])
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, policy, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=300) for _ in procs), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    covers = np.concatenate([r[1] for r in res], axis=1)
    np.testing.assert_array_equal(covers[0], ref.grid[0, 1])
    np.testing.assert_array_equal(covers[1], ref.grid[0, 2])
    for r in res:                                   # agents and counters are replicated: every rank must agree
        np.testing.assert_array_equal(r[2][0], ref.agent_indices[0])
        np.testing.assert_array_equal(r[2][1], ref.agent_states[0, :, 0])
        assert r[3][0] == int(done_at[0])
        np.testing.assert_array_equal(r[3][1], ada[0, :, 0])
        assert r[4] == STEPS and r[5] == ref.L


@pytest.mark.parametrize("world_size", [1, 2])
def test_thread_bands_match_full_torus_oracle(world_size):
    from band_helpers import OracleBand, ThreadComm, make_state, run_threads
    from therldaisyworld_b200.banded import BandedDaisyWorld
    (light, dark, ai, st), ref, done_at, ada = _reference("antigreedy", None)
    shared = ThreadComm.Shared(world_size)
    worlds = [BandedDaisyWorld(N, N_AGENTS, rank=r, world_size=world_size, band_factory=OracleBand,
                               comm=ThreadComm(r, world_size, shared) if world_size > 1 else None) for r in range(world_size)]

    def go(w):
        w.load_state(light, dark, ai, st)
        w.run(STEPS, "antigreedy", chunk=7)

    run_threads(worlds, go)
    covers = np.concatenate([w.local_covers() for w in worlds], axis=1)
    for w in worlds:
        _check(w, ref, done_at, ada, covers)


def test_band_rows_partition():
    from therldaisyworld_b200.banded import band_rows
    assert [band_rows(512, 4, r) for r in range(4)] == [(0, 128), (128, 256), (256, 384), (384, 512)]
    with pytest.raises(ValueError):
        band_rows(192, 2, 0)
