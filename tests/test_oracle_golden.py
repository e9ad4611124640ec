"""Pin the NumPy oracle against trajectories recorded from the live reference.

The fixtures in tests/golden were produced by oracle/gen_golden.py from the unmodified
reference (daisy/daisy_world_rl.py + daisy/agents/greedy.py).  Bar: every recorded quantity is
reproduced exactly (integers, bools) or value-identical in fp64 (==, so -0.0 == 0.0)."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR, golden_names
from helpers import consume_policy_draws, restore_stream, uses_global_stream
from oracle.daisy_numpy import OracleDaisyWorld, OracleGreedy, OracleMLP, env_from_golden, neighborhood_mask


def replay(z, env, check_every_step=True):
    meta = json.loads(str(z["meta"]))
    T = meta["steps"]
    ck = {int(s): i for i, s in enumerate(z["ckpt_steps"])}
    B, n = meta["B"], meta["n"]
    done_at = np.zeros((B,), dtype=np.int64)
    agents_done_at = np.zeros((B, n, 1), dtype=np.int64)
    stream = uses_global_stream(z, meta)
    if stream:
        restore_stream(z)
    for t in range(T):
        if stream:
            consume_policy_draws(z, meta, t)
        a = z["actions"][t]
        action = None if (a.shape == (1, 1, 1) and a[0, 0, 0] == -1) else a
        assert env.L == z["L"][t]
        obs, reward, done, info = env.step(action)
        assert info == {}
        np.testing.assert_array_equal(env.agent_indices, z["agent_indices"][t])
        np.testing.assert_array_equal(env.agent_states, z["agent_states"][t])
        np.testing.assert_array_equal(reward, z["reward"][t])
        np.testing.assert_array_equal(done, z["done"][t])
        np.testing.assert_array_equal(env.grid.sum(axis=(-2, -1)), z["chan_sum"][t])
        if (t + 1) in ck:
            i = ck[t + 1]
            np.testing.assert_array_equal(env.grid, z["ckpt_grid"][i])
            np.testing.assert_array_equal(obs, z["ckpt_obs"][i])
        grid_done = env.grid[:, 1:3].max(axis=(1, 2, 3)) <= 0.005
        done_at += 1 - 1 * grid_done
        if n:
            agents_done_at += 1 - 1 * done
    assert env.L == z["L"][T]
    assert env.step_count == meta["final_step_count"]
    np.testing.assert_array_equal(done_at, z["done_at"])
    np.testing.assert_array_equal(agents_done_at, z["agents_done_at"])
    return env


@pytest.mark.parametrize("name", golden_names())
def test_oracle_replays_reference_trajectory(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    env, meta = env_from_golden(z)
    np.testing.assert_array_equal(env.get_obs(env.agent_indices), z["init_obs"])
    env = replay(z, env)
    # unrounded side-effect diagnostics of the last forward: FFT round-off in the reference
    # (|delta| <= 2.2e-15 on convolutions) => 1e-9 relative (north_star tolerance)
    for key, attr in [("diag_temp", "temp"), ("diag_temp_light", "temp_light"), ("diag_temp_dark", "temp_dark"),
                      ("diag_temp_effective", "temp_effective"), ("diag_dead_temp", "dead_temp"),
                      ("diag_beta", "beta"), ("diag_beta_l", "beta_l"), ("diag_beta_d", "beta_d")]:
        np.testing.assert_allclose(getattr(env, attr), z[key], rtol=1e-9, atol=1e-12, err_msg=key)
    np.testing.assert_allclose(env.growth, z["diag_growth"], rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("name", ["greedy_n16_b4_todeath", "antigreedy_n8_b16_todeath", "greedy_n5_b3_n2_todeath",
                                  "greedy_n64_b2_120", "greedy_n17_b2_params_200"])
def test_oracle_greedy_policy_reproduces_reference_actions(name):
    """OracleGreedy, driven by oracle observations, picks the actions the reference Greedy picked."""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    env, meta = env_from_golden(z)
    agent = OracleGreedy(epsilon=0.0, greedy=meta["policy"]["kind"] != "antigreedy")
    obs = env.get_obs(env.agent_indices)
    for t in range(meta["steps"]):
        action = agent(obs)
        np.testing.assert_array_equal(action, z["actions"][t])
        obs, _, _, _ = env.step(action)
    np.testing.assert_array_equal(env.grid, z["ckpt_grid"][-1])


@pytest.mark.parametrize("name", ["mlp_n16_b4_200", "mlp_n64_b2_n6_60", "mlp_n16_b4_mixed_150", "mlp_moore_n16_b3_100",
                                  "mlp_circular_n16_b2_60"])
def test_oracle_mlp_policy_reproduces_reference_actions(name):
    """OracleMLP with the recorded weights, driven by oracle observations, picks the actions the reference MLP picked."""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    env, meta = env_from_golden(z)
    agent = OracleMLP(z["mlp_params"])
    obs = env.get_obs(env.agent_indices)
    for t in range(meta["steps"]):
        action = agent(obs)
        np.testing.assert_array_equal(action, z["actions"][t])
        obs, _, _, _ = env.step(action)
    np.testing.assert_array_equal(env.grid, z["ckpt_grid"][-1])


@pytest.mark.parametrize("name", ["cfg1_n16_b1_noagents_todeath", "greedy_n16_b4_todeath", "randint_n7_b3_n16_200",
                                  "greedy_n17_b2_params_200"])
def test_oracle_reset_reproduces_reference_rng_order(name):
    """Constructor + reset consume the global legacy RNG exactly like the reference (A14)."""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    meta = json.loads(str(z["meta"]))
    np.random.seed(meta["seed"])
    env = OracleDaisyWorld(**meta["ctor"])
    for k, v in meta["attrs"].items():
        env.set_use_microclimate(v) if k == "use_microclimate" else setattr(env, k, v)
    obs = env.reset()
    np.testing.assert_array_equal(env.agent_indices, z["init_agent_indices"])
    np.testing.assert_array_equal(env.agent_states, z["init_agent_states"])
    np.testing.assert_array_equal(env.grid[:, :3], z["init_grid"][:, :3])
    # initial temperatures are unrounded: FFT-vs-stencil round-off only
    np.testing.assert_allclose(env.grid[:, 3:], z["init_grid"][:, 3:], rtol=1e-12)
    np.testing.assert_allclose(obs, z["init_obs"], rtol=1e-12)


def test_fft_mode_matches_stencil_mode():
    z = np.load(os.path.join(GOLDEN_DIR, "greedy_n16_b4_todeath.npz"))
    env, _ = env_from_golden(z, conv="fft")
    replay(z, env)


def test_neighborhood_masks():
    """Mirrors the reference's tests/daisy/test_functional.py:17-44."""
    for kr in (1, 2, 3):
        for mode in ("moore", "von_neumann", "circular", "nonsense"):
            m = neighborhood_mask(kr, mode)
            assert m.shape == (2 * kr + 1, 2 * kr + 1)
            assert m[kr, kr] == 1.0
            assert m[0, 0] == (1.0 if mode == "moore" else 0.0)


def test_neighborhood_masks_equal_the_reference_recording():
    """A15: the oracle's and the PRODUCT's make_neighborhood against masks recorded from the live reference
    (daisy/nn/functional.py:51-103; oracle/gen_golden.py::record_masks), every mode, radius 1..3."""
    from therldaisyworld_b200.env import make_neighborhood
    z = np.load(os.path.join(GOLDEN_DIR, "ref_masks.npz"))
    for kr in (1, 2, 3):
        for mode in ("moore", "von_neumann", "circular", "nonsense"):
            want = z[f"{mode}_{kr}"]
            np.testing.assert_array_equal(neighborhood_mask(kr, mode), want)
            got = make_neighborhood(kr, mode)
            assert got.dtype == want.dtype and got.shape == want.shape
            np.testing.assert_array_equal(got, want)
    np.testing.assert_array_equal(make_neighborhood(), z["default"])
