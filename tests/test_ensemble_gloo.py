"""Host-side logic of the sharded ensemble (therldaisyworld_b200/ensemble.py) on CPU: two gloo ranks, each with an
oracle-backed shard, must reproduce the single-process lifespan experiment recorded from the reference (same stopping
step, same lifespans, same statistics)."""
import json
import os
import socket
import sys

import numpy as np
import pytest

from conftest import GOLDEN_DIR, ROOT


class OracleShard:
    """Stand-in for DeviceShard backed by the C oracle (test infrastructure)."""

    def __init__(self, env_like, grid, agent_indices, agent_states):
        from oracle.daisy_c import COracleWorld
        self.w = COracleWorld(env_like, grid=grid, agent_indices=agent_indices, agent_states=agent_states)
        self.done_at = np.zeros((self.w.B,), dtype=np.int64)
        self.agents_done_at = np.zeros((self.w.B, self.w.n, 1), dtype=np.int64)
        self._ck = None

    def begin(self):
        self.done_at[:] = 0
        self.agents_done_at[:] = 0

    def checkpoint_save(self):
        import copy
        self._ck = (self.w.grid.copy(), self.w.agent_indices.copy(), self.w.agent_states.copy(), copy.copy(self.w.clk),
                    self.done_at.copy(), self.agents_done_at.copy())

    def checkpoint_restore(self):
        import copy
        g, ai, st, clk, d, a = self._ck
        self.w.grid[:] = g; self.w.agent_indices[:] = ai; self.w.agent_states[:] = st; self.w.clk = copy.copy(clk)
        self.done_at[:] = d; self.agents_done_at[:] = a

    def run_chunk(self, K, policy, actions=None, seed=0, masked=False):
        mask = 0
        if masked:
            self._alive_bits = np.zeros(self.agents_done_at.shape, dtype=np.uint64)
        for j in range(K):
            steps, d, a = self.w.run(1, policy if policy != "replay" else "replay",
                                     actions=None if actions is None else actions[j:j + 1])
            self.done_at += d
            self.agents_done_at += a
            if masked:                     # what dw_run_chunk_masked records: bit j = agent not done after step j
                self._alive_bits |= (a.astype(np.uint64) << np.uint64(j))
            if (d == 0).all():
                mask |= 1 << j
        return mask

    def stats(self, like):
        d = self.done_at.astype(np.float64); a = self.agents_done_at.astype(np.float64)
        like[:] = like.new_tensor([d.size, d.sum(), (d * d).sum(), a.size, a.sum(), (a * a).sum(), 0.0, 0.0])
        return like


class TrimmingOracleShard(OracleShard):
    """OracleShard that also offers the statistics-only protocol of DeviceShard (trim_supported / masked run_chunk / trim):
    from the second segment on, like a device shard whose first step is literal. rank-dependent support exercises the
    per-segment agreement between the ranks (a rank that cannot trim forces the checkpoint path on everyone)."""

    def __init__(self, *a, supports=True, **k):
        super().__init__(*a, **k)
        self.supports = supports
        self.segments = 0
        self.trims = 0
        self.checkpoints = 0

    def checkpoint_save(self):
        self.checkpoints += 1
        super().checkpoint_save()

    def trim_supported(self, policy):
        return self.supports and self.segments > 0

    def suggest_segment(self):
        return 32

    def run_chunk(self, K, policy, actions=None, seed=0, masked=False):
        self.segments += 1
        return super().run_chunk(K, policy, actions, seed, masked=masked)

    def trim(self, j):
        self.trims += 1
        surplus = np.zeros(self.agents_done_at.shape, dtype=np.int64)
        for t in range(j + 1, 64):
            surplus += ((self._alive_bits >> np.uint64(t)) & np.uint64(1)).astype(np.int64)
        self.agents_done_at -= surplus



def _worker_trim(rank, world, port, name, q, supports):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from oracle.daisy_numpy import env_from_golden
    from therldaisyworld_b200.ensemble import shard_range, simulate_lifespan
    dist.init_process_group("gloo", rank=rank, world_size=world)
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    env, meta = env_from_golden(z)
    lo, hi = shard_range(meta["B"], world, rank)
    shard = TrimmingOracleShard(env, z["init_grid"][lo:hi], z["init_agent_indices"][lo:hi], z["init_agent_states"][lo:hi],
                                supports=supports[rank])
    policy = {"antigreedy": "antigreedy", "greedy": "greedy", "random": "replay"}[meta["policy"]["kind"]]
    actions = z["actions"][:, lo:hi] if policy == "replay" else None
    out = simulate_lifespan(shard, policy=policy, actions=actions)
    q.put((rank, lo, hi, out, shard.done_at.copy(), shard.agents_done_at.copy(), shard.trims, shard.checkpoints, shard.segments))
    dist.destroy_process_group()


def _worker(rank, world, port, name, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from oracle.daisy_numpy import env_from_golden
    from therldaisyworld_b200.ensemble import shard_range, simulate_lifespan
    dist.init_process_group("gloo", rank=rank, world_size=world)
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    env, meta = env_from_golden(z)
    lo, hi = shard_range(meta["B"], world, rank)
    shard = OracleShard(env, z["init_grid"][lo:hi], z["init_agent_indices"][lo:hi], z["init_agent_states"][lo:hi])
    policy = {"antigreedy": "antigreedy", "greedy": "greedy", "random": "replay"}[meta["policy"]["kind"]]
    actions = z["actions"][:, lo:hi] if policy == "replay" else None
    out = simulate_lifespan(shard, policy=policy, actions=actions, segment=64)
    q.put((rank, lo, hi, out, shard.done_at.copy(), shard.agents_done_at.copy()))
    dist.destroy_process_group()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


@pytest.mark.parametrize("name", ["antigreedy_n8_b16_todeath", "random_n8_b8_todeath"])
def test_two_rank_ensemble_matches_reference(name):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, name, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    meta = json.loads(str(z["meta"]))
    done_at = np.concatenate([r[4] for r in res])
    agents = np.concatenate([r[5] for r in res])
    np.testing.assert_array_equal(done_at, z["done_at"])
    np.testing.assert_array_equal(agents, z["agents_done_at"])
    for r in res:
        out = r[3]
        assert out["steps"] == meta["steps"] and out["all_done"] and out["worlds"] == meta["B"]
        assert out["biosphere_lifespan_mean"] == pytest.approx(z["done_at"].mean(), rel=1e-12)
        assert out["biosphere_lifespan_sem"] == pytest.approx(z["done_at"].std() / np.sqrt(meta["B"]), rel=1e-9)
        assert out["agent_lifespan_mean"] == pytest.approx(z["agents_done_at"].mean(), rel=1e-12)
        assert out["agent_lifespan_sem"] == pytest.approx(z["agents_done_at"].std() / np.sqrt(meta["B"]), rel=1e-9)


@pytest.mark.parametrize("supports", [(True, True), (True, False)])
def test_two_rank_ensemble_with_trimmed_segments(supports):
    """Statistics-only protocol (no rewinding): both ranks trim -> one checkpoint (the first segment) and one trim each; one
    rank cannot -> every segment is checkpointed on both ranks and the stopping step is reached by rewind + replay. Same
    lifespans as the reference recording either way."""
    import torch.multiprocessing as mp
    name = "antigreedy_n8_b16_todeath"
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_trim, args=(r, 2, port, name, q, supports)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    meta = json.loads(str(z["meta"]))
    np.testing.assert_array_equal(np.concatenate([r[4] for r in res]), z["done_at"])
    np.testing.assert_array_equal(np.concatenate([r[5] for r in res]), z["agents_done_at"])
    for r in res:
        assert r[3]["steps"] == meta["steps"] and r[3]["all_done"]
        trims, checkpoints, segments = r[6], r[7], r[8]
        if all(supports):
            assert trims == 1 and checkpoints == 1 and segments >= 2
        else:
            assert trims == 0 and checkpoints >= 2


def test_shard_range_partitions():
    from therldaisyworld_b200.ensemble import shard_range
    for total in (1, 7, 1000, 1_000_003):
        for ws in (1, 2, 3, 8):
            spans = [shard_range(total, ws, r) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
