"""GPU parity of dw_run (K fused steps, on-device policy, lifespan counters) against the reference-recorded
fixtures and the C oracle.  Integer lifespans must be identical; grids value-identical."""
import numpy as np
import pytest

from helpers import load_golden, product_env_from_golden

pytestmark = pytest.mark.gpu

CASES = [("greedy_n16_b4_todeath", "greedy"), ("antigreedy_n8_b16_todeath", "antigreedy"),
         ("random_n8_b8_todeath", "replay"), ("greedy_n5_b3_n2_todeath", "greedy"),
         ("cfg1_n16_b1_noagents_todeath", "none"), ("halfrandom_n16_b4_300", "replay"),
         ("greedy_n64_b2_120", "greedy"), ("neutral_antigreedy_n64_b1_40", "antigreedy"),
         ("randint_n7_b3_n16_200", "replay"), ("greedy_n17_b2_params_200", "greedy"),
         ("none_n16_b2_n4_40", "none"), ("rampupdown_n8_b2_100", "greedy")]


@pytest.mark.parametrize("name,policy", CASES)
def test_run_reproduces_reference_lifespans_and_final_state(name, policy):
    z, meta = load_golden(name)
    env = product_env_from_golden(z, meta)
    env.reset_lifespans()
    K = meta["steps"]
    actions = z["actions"] if policy == "replay" else None
    steps, alive, hit = env.run(100000 if meta["to_death"] and policy != "replay" else K, policy=policy, actions=actions,
                                stop_all_done=meta["to_death"])
    assert steps == K
    if meta["to_death"]:
        assert hit and alive == 0
    done_at, agents_done_at = env.lifespans()
    np.testing.assert_array_equal(done_at, z["done_at"])
    np.testing.assert_array_equal(agents_done_at, z["agents_done_at"])
    np.testing.assert_array_equal(env.agent_indices, z["agent_indices"][-1])
    np.testing.assert_array_equal(env.agent_states, z["agent_states"][-1])
    np.testing.assert_array_equal(env.grid, z["ckpt_grid"][-1])
    np.testing.assert_array_equal(env.get_obs(env.agent_indices), z["ckpt_obs"][-1])
    assert env.L == z["L"][K] and env.step_count == meta["final_step_count"]
    np.testing.assert_allclose(env.temp, z["diag_temp"], rtol=1e-9)
    np.testing.assert_allclose(env.growth, z["diag_growth"], rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("name,policy", [("greedy_n16_b4_todeath", "greedy"), ("greedy_n64_b2_120", "greedy"),
                                         ("random_n8_b8_todeath", "replay")])
def test_run_in_pieces_equals_single_steps(name, policy):
    """Chunk boundaries and mixing run()/step() must not change anything: compare checkpoints on the way."""
    z, meta = load_golden(name)
    env = product_env_from_golden(z, meta)
    t = 0
    for s, i in sorted((int(s), i) for i, s in enumerate(z["ckpt_steps"])):
        k = s - t
        if k:
            acts = z["actions"][t:s] if policy == "replay" else None
            env.run(k, policy=policy, actions=acts)
            t = s
        np.testing.assert_array_equal(env.grid, z["ckpt_grid"][i])
        np.testing.assert_array_equal(env.get_obs(env.agent_indices), z["ckpt_obs"][i])
        if t < meta["steps"]:   # interleave one materialising step
            a = z["actions"][t]
            obs, reward, done, _ = env.step(None if a[0, 0, 0] == -1 else a)
            np.testing.assert_array_equal(reward, z["reward"][t])
            t += 1


def test_run_large_ensemble_matches_c_oracle():
    """BASELINE config 2 shape at reduced batch (64 worlds, 64x64, greedy) against the C oracle, full life."""
    from therldaisyworld_b200 import RLDaisyWorld
    from oracle.daisy_c import COracleWorld
    np.random.seed(13)
    env = RLDaisyWorld(grid_dimension=64)
    env.batch_size = 64
    env.reset()
    ref = COracleWorld(env)
    done_at, agents_done_at = env.simulate_lifespan(policy="greedy")
    steps, d2, a2 = ref.run(100000, "greedy", stop_all_done=True)
    assert env.step_count == steps
    np.testing.assert_array_equal(done_at, d2)
    np.testing.assert_array_equal(agents_done_at, a2)
    np.testing.assert_array_equal(env.grid, ref.grid)


def test_cfg2_full_size_lifespans_identical_to_reference():
    """BASELINE config 2 at full size through the CUDA path: lifespans and final-state checksums identical to
    the live reference's (fixture recorded by oracle/gen_golden_cfg2.py)."""
    from therldaisyworld_b200 import RLDaisyWorld
    z, meta = load_golden("big_cfg2_greedy_n64_b1000")
    np.random.seed(13)
    env = RLDaisyWorld(grid_dimension=64)
    env.batch_size = 1000
    env.reset()
    np.testing.assert_array_equal(env.agent_indices, z["init_agent_indices"])
    done_at, agents_done_at = env.simulate_lifespan(policy="greedy")
    assert env.step_count == meta["steps"]
    np.testing.assert_array_equal(done_at, z["done_at"])
    np.testing.assert_array_equal(agents_done_at, z["agents_done_at"])
    np.testing.assert_array_equal(env.grid.sum(axis=(-2, -1)), z["final_chan_sum"])
    np.testing.assert_array_equal(env.agent_states, z["final_agent_states"])
    np.testing.assert_array_equal(env.agent_indices, z["final_agent_indices"])


def test_device_shard_ensemble_single_rank():
    """ensemble.simulate_lifespan over a DeviceShard == the reference's recorded experiment (statistics included)."""
    from therldaisyworld_b200.ensemble import DeviceShard, simulate_lifespan
    z, meta = load_golden("antigreedy_n8_b16_todeath")
    env = product_env_from_golden(z, meta)
    out = simulate_lifespan(DeviceShard(env), policy="antigreedy", device="cuda")
    assert out["steps"] == meta["steps"] and out["all_done"]
    done_at, agents_done_at = env.lifespans()
    np.testing.assert_array_equal(done_at, z["done_at"])
    np.testing.assert_array_equal(agents_done_at, z["agents_done_at"])
    assert out["biosphere_lifespan_mean"] == pytest.approx(z["done_at"].mean(), rel=1e-12)
    assert out["agent_lifespan_sem"] == pytest.approx(z["agents_done_at"].std() / 4.0, rel=1e-9)


def test_device_side_reset_distribution_and_run():
    """dw_init_random: same distribution as the reference's reset (different RNG stream); runs through the fused path
    with the device-side random policy, and the run is reproducible for a fixed seed."""
    from therldaisyworld_b200 import RLDaisyWorld
    np.random.seed(0)
    env = RLDaisyWorld(grid_dimension=64)
    env.batch_size = 64
    outs = []
    for rep in range(2):
        env.reset_on_device(seed=7)
        g = env.grid
        for ch, prop, init in ((1, env.light_proportion, env.initial_al), (2, env.dark_proportion, env.initial_ad)):
            x = g[:, ch]
            assert abs((x > 0).mean() - prop) < 0.01
            assert x.max() < init and abs(x[x > 0].mean() - init / 2) < 0.002
        np.testing.assert_array_equal(g[:, 0], (env.p - g[:, 1]) - g[:, 2])
        assert (g[:, 3:6] > 200).all() and (env.agent_states == 1).all()
        assert env.agent_indices.min() >= 0 and env.agent_indices.max() < 64
        env.reset_lifespans()
        env.run(200, policy="random", seed=11)
        outs.append((env.grid.copy(), env.agent_states.copy(), env.lifespans()[1].copy()))
    np.testing.assert_array_equal(outs[0][0], outs[1][0])
    np.testing.assert_array_equal(outs[0][1], outs[1][1])
    np.testing.assert_array_equal(outs[0][2], outs[1][2])


@pytest.mark.parametrize("N,sizes", [(64, (16, 16, 16)), (8, (70, 37, 100)), (16, (19, 30, 5)), (32, (7, 2, 9)), (96, (3, 2, 4))])
def test_ensemble_sharding_is_invisible(N, sizes):
    """An ensemble run as one handle or as three shards with world offsets (the multi-rank layout) gives the same worlds:
    device reset and device random policy are keyed by the GLOBAL world index. For worlds below 64x64 the shards also
    regroup the worlds that share a CTA (k_fused_sub64_persist), which must not matter either."""
    from therldaisyworld_b200 import RLDaisyWorld

    def make(B, offset):
        np.random.seed(1)
        env = RLDaisyWorld(grid_dimension=N)
        env.batch_size = B
        env.reset_on_device(seed=5, world_offset=offset)
        env.reset_lifespans()
        env.run(150, policy="random", seed=3)
        return env.grid.copy(), env.agent_states.copy(), env.lifespans()

    whole = make(sum(sizes), 0)
    parts = [make(b, int(off)) for b, off in zip(sizes, np.cumsum((0,) + sizes[:-1]))]
    np.testing.assert_array_equal(np.concatenate([p[0] for p in parts]), whole[0])
    np.testing.assert_array_equal(np.concatenate([p[1] for p in parts]), whole[1])
    np.testing.assert_array_equal(np.concatenate([p[2][1] for p in parts]), whole[2][1])


@pytest.mark.parametrize("N,n,policy", [(8, 4, "greedy"), (16, 9, "antigreedy"), (32, 4, "random"), (64, 4, "greedy"), (96, 5, "greedy"),
                                        (33, 3, "greedy"), (128, 6, "random"), (14, 3, "greedy"), (75, 4, "antigreedy")])
def test_torus_translation_invariance(N, n, policy):
    """A size-independent property of the path: the world is a torus and nothing in the step depends on absolute
    coordinates, so shifting the initial covers and agents by (sx, sy) must shift the whole run by (sx, sy) -- through
    every fused kernel (several worlds per CTA, 64x64, 4x4 tiles, one cell per thread) and the materialising steps.
    Exercises every wrap-around (rows, halo columns, agents crossing the seam) without the oracle."""
    from therldaisyworld_b200 import RLDaisyWorld
    B, sx, sy = 5, N - 3, 5 % N
    np.random.seed(N)
    a = RLDaisyWorld(grid_dimension=N, n_agents=n)
    a.batch_size = B
    a.reset()
    b = RLDaisyWorld(grid_dimension=N, n_agents=n)
    b.batch_size = B
    b.reset()
    g0, ai0, st0 = a.grid.copy(), a.agent_indices.copy(), a.agent_states.copy()
    b.grid = np.roll(g0, (sx, sy), axis=(2, 3))
    b.agent_indices = (ai0 + np.array([sx, sy])) % N
    b.agent_states = st0.copy()
    for env in (a, b):
        env.reset_lifespans()
        env.run(1, policy=policy, seed=4)          # literal first step from the off-lattice reset state
        env.run(90, policy=policy, seed=4)         # fused
        env.step_policy(policy, seed=4)            # materialising
    np.testing.assert_array_equal(np.roll(a.grid, (sx, sy), axis=(2, 3)), b.grid)
    np.testing.assert_array_equal((a.agent_indices + np.array([sx, sy])) % N, b.agent_indices)
    np.testing.assert_array_equal(a.agent_states, b.agent_states)
    np.testing.assert_array_equal(a.lifespans()[0], b.lifespans()[0])
    np.testing.assert_array_equal(a.lifespans()[1], b.lifespans()[1])


@pytest.mark.parametrize("N,B,policy", [(64, 40, "greedy"), (64, 25, "random"), (16, 70, "greedy"), (32, 9, "antigreedy")])
def test_trimmed_lifespans_equal_checkpoint_and_replay(monkeypatch, N, B, policy):
    """ensemble.simulate_lifespan, statistics only: segments run 'masked' past the stopping step and the surplus is trimmed out
    of the agents' counters (dw_run_chunk_masked / dw_trim_lifespans) -- against the checkpoint + rewind + replay path
    (DW_NO_TRIM=1): identical lifespan counters, steps and statistics."""
    from therldaisyworld_b200 import RLDaisyWorld
    from therldaisyworld_b200.ensemble import DeviceShard, simulate_lifespan
    res = []
    for no_trim in (False, True):
        if no_trim:
            monkeypatch.setenv("DW_NO_TRIM", "1")
        else:
            monkeypatch.delenv("DW_NO_TRIM", raising=False)
        np.random.seed(3)
        env = RLDaisyWorld(grid_dimension=N)
        env.batch_size = B
        env.reset_on_device(seed=11)
        shard = DeviceShard(env)
        out = simulate_lifespan(shard, policy=policy, seed=5, device="cuda")
        if not no_trim:
            assert shard.trim_supported(policy)
        res.append((out, env.lifespans()))
    (a, la), (b, lb) = res
    assert a == b and a["all_done"] and a["steps"] % 64 != 0
    np.testing.assert_array_equal(la[0], lb[0])
    np.testing.assert_array_equal(la[1], lb[1])
