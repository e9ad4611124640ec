"""Next-row N3 (diagnostics feed): ensemble statistics of the unrounded diagnostic fields and of the covers, reduced on the
device, against NumPy on the downloaded fields -- after a reset, after single steps and after fused runs."""
import numpy as np
import pytest

from helpers import load_golden, product_env_from_golden

pytestmark = pytest.mark.gpu


def _check(env):
    for name in ("temp", "temp_light", "beta_l", "growth"):
        s, f = env.diag_stats(name), getattr(env, name)
        np.testing.assert_allclose([s["mean"], s["min"], s["max"]], [f.mean(), f.min(), f.max()], rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(s["std"], f.std(), rtol=1e-9)        # NumPy's own two-pass std vs the shifted one-pass sum
    c, g = env.cover_stats(), env.grid
    np.testing.assert_allclose([c["mean_light"], c["mean_dark"], c["max_light"], c["max_dark"]],
                               [g[:, 1].mean(), g[:, 2].mean(), g[:, 1].max(), g[:, 2].max()], rtol=1e-12)


def test_device_statistics_match_numpy_in_every_state():
    z, meta = load_golden("greedy_n64_b2_120")
    env = product_env_from_golden(z, meta)
    env.run(1, policy="greedy")          # lean first step: pre-state = cover planes
    stats_before = env.cover_stats()     # lattice path (no grid materialised yet)
    _check(env)
    assert stats_before == env.cover_stats()
    env.run(40, policy="greedy")         # fused: pre-state = lattice
    _check(env)
    env.step_policy("greedy")            # materialising step: pre-state = grid
    _check(env)
    # the reference's global mean temperature of the last forward (notebook_helpers.py:50)
    np.testing.assert_allclose(env.diag_stats("temp")["mean"], env.temp.mean(), rtol=1e-12)


def test_series_helper_records_the_plot_feed():
    z, meta = load_golden("greedy_n16_b4_todeath")
    env = product_env_from_golden(z, meta)
    series = env.run_with_series(96, every=32)
    assert [s["step"] for s in series] == [32, 64, 96]
    assert all(200 < s["temp_mean"] < 400 and 0 <= s["light"] <= 1 for s in series)
    np.testing.assert_allclose(series[-1]["temp_mean"], env.temp.mean(), rtol=1e-12)


def test_in_kernel_series_matches_oracle_per_step():
    """dw_run_series: the per-step global mean temperature (env.temp.mean() of the reference, notebook_helpers.py:50) and
    mean covers, reduced inside the fused kernel, against the NumPy oracle stepped with its Greedy restatement."""
    from oracle.daisy_numpy import OracleGreedy, env_from_golden
    z, meta = load_golden("greedy_n64_b2_120")
    env = product_env_from_golden(z, meta)
    K = 60
    series = env.run_series(K, policy="greedy")
    oenv, _ = env_from_golden(z)
    agent = OracleGreedy()
    obs = oenv.get_obs(oenv.agent_indices)
    ref = np.zeros((K, 3))
    for t in range(K):
        obs, _, _, _ = oenv.step(agent(obs))
        ref[t] = [oenv.temp.mean(), oenv.grid[:, 1].mean(), oenv.grid[:, 2].mean()]
    np.testing.assert_allclose(series[:, 0], ref[:, 0], rtol=1e-9)          # north_star tolerance for fp64 fields
    np.testing.assert_allclose(series[:, 1:], ref[:, 1:], rtol=1e-12)
    np.testing.assert_array_equal(env.grid, oenv.grid)                      # the run itself is unchanged by the series mode
    # a second call continues from the lattice state (all steps inside the kernel)
    s2 = env.run_series(10, policy="greedy")
    for t in range(10):
        obs, _, _, _ = oenv.step(agent(obs))
        np.testing.assert_allclose(s2[t], [oenv.temp.mean(), oenv.grid[:, 1].mean(), oenv.grid[:, 2].mean()], rtol=1e-9)


@pytest.mark.parametrize("name,policy", [("greedy_n16_b4_todeath", "greedy"), ("antigreedy_n8_b16_todeath", "antigreedy"),
                                         ("greedy_n17_b2_params_200", "greedy")])
def test_series_for_other_world_sizes_matches_oracle_per_step(name, policy):
    """run_series on the other shapes: 16x16 and 8x8 (in-kernel series mode of the sub-64 persistent kernel; the 8x8 case fills
    16 of the 64 world slots of a CTA) and 17x17 (sampled between one-step launches): same per-step means as the oracle."""
    from oracle.daisy_numpy import OracleGreedy, env_from_golden
    z, meta = load_golden(name)
    env = product_env_from_golden(z, meta)
    K = 40
    series = env.run_series(K, policy=policy)
    oenv, _ = env_from_golden(z)
    agent = OracleGreedy(greedy=policy == "greedy")
    obs = oenv.get_obs(oenv.agent_indices)
    for t in range(K):
        obs, _, _, _ = oenv.step(agent(obs))
        np.testing.assert_allclose(series[t], [oenv.temp.mean(), oenv.grid[:, 1].mean(), oenv.grid[:, 2].mean()], rtol=1e-9)
    np.testing.assert_array_equal(env.grid, oenv.grid)


def test_fp32_export_within_stated_tolerance(monkeypatch):
    """fp32 export (states that are not lattice-resident, or DW_F32_EXPORT_ONLY=1): one rounding of the exact fp64 fields, so
    6e-8 holds (BASELINE.json's tolerance for fp32 is 1e-5 relative)."""
    monkeypatch.setenv("DW_F32_EXPORT_ONLY", "1")
    z, meta = load_golden("greedy_n64_b2_120")
    env = product_env_from_golden(z, meta)
    env.run(30, policy="greedy")
    g32, o32 = env.grid_f32(), env.observe_f32()
    assert g32.dtype == np.float32 and o32.dtype == np.float32
    np.testing.assert_allclose(g32, env.grid, rtol=1e-5, atol=0)
    np.testing.assert_allclose(g32, env.grid, rtol=6e-8, atol=0)
    np.testing.assert_allclose(o32, env.observe(), rtol=6e-8, atol=0)
    assert env.f32_stats()["cells"] == 0


def _check_f32_against_f64(env, tol_T=1e-5):
    """fp32-arithmetic materialisation (k_forward_f32) against the bit-exact fp64 grid of the same state."""
    before = env.f32_stats()
    g32 = env.grid_f32()                 # first: reading env.grid hands the state to the host mirror
    after = env.f32_stats()
    g64 = env.grid
    assert g32.dtype == np.float32
    assert after["cells"] - before["cells"] == g64.shape[0] * g64.shape[2] * g64.shape[3], "the fp32-arithmetic path did not run"
    # covers and bare fraction: exactly the binary32 rounding of the reference values (b' through the screened tiers)
    for ch in (0, 1, 2):
        np.testing.assert_array_equal(g32[:, ch], g64[:, ch].astype(np.float32))
    np.testing.assert_array_equal(g32[:, 6], 0)
    # channel 4 carries the agent stamp at the agents' cells
    B, n = env.agent_indices.shape[:2]
    stamp = np.zeros(g64[:, 4].shape, dtype=bool)
    for b in range(B):
        for i in range(n):
            x, y = env.agent_indices[b, i]
            stamp[b, x, y] = True
    np.testing.assert_array_equal(g32[:, 4][stamp], g64[:, 4][stamp].astype(np.float32))
    # temperatures: fp32 arithmetic, the north_star's fp32 tolerance (1e-5 relative); observed ~3e-6 (a 0.001 K rounding flip)
    for ch in (3, 4, 5):
        m = ~stamp if ch == 4 else np.ones_like(stamp)
        rel = np.abs(g32[:, ch][m].astype(np.float64) - g64[:, ch][m]) / np.abs(g64[:, ch][m])
        assert rel.max() <= tol_T, (ch, rel.max())
    return after, g32


def test_fp32_mode_arithmetic_within_1e5_of_the_reference_fields():
    """north_star fp32 mode: after fused steps the 7-channel grid is materialised with fp32 ARITHMETIC from the packed lattice
    (csrc/dw_f32.cuh). env.grid here is bit-equal to the recorded reference grid (test_run_* pin that), so comparing against
    it compares against the reference: covers / bare fraction exact (as binary32), temperatures within 1e-5 relative."""
    for name, K in (("greedy_n64_b2_120", 30), ("greedy_n16_b4_todeath", 150), ("greedy_n17_b2_params_200", 60),
                    ("neutral_antigreedy_n64_b1_40", 25), ("random_n8_b8_todeath", 90), ("greedy_moore_n64_b2_n6_60", 45)):
        z, meta = load_golden(name)
        env = product_env_from_golden(z, meta)
        pol = {"greedy": "greedy", "antigreedy": "antigreedy"}.get(meta["policy"]["kind"], "random")
        env.run(K, policy=pol, seed=3)
        assert env.residency()["lattice"], name
        o32 = env.observe_f32()
        st, g32 = _check_f32_against_f64(env)
        assert st["fp64_tier"] <= 0.2 * st["cells"], (name, st)          # the fp32 evaluation decides the large majority of cells
        # the observation windows cut from the fp32 grid
        o64 = env.observe()
        assert o32.dtype == np.float32
        np.testing.assert_allclose(o32, o64, rtol=1e-5, atol=0)
        np.testing.assert_array_equal(o32[:, :, 0:3], o64[:, :, 0:3].astype(np.float32))


def test_fp32_mode_random_physics():
    """Random constants (albedos, S, g, gamma, dt, T_opt, luminosity ramp, microclimate off) through the fp32-arithmetic
    materialisation: the screen of the bare fraction (error bound from the handle's constants) must keep channel 0 exact."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    from fuzz_fast_path import draw
    from therldaisyworld_b200 import RLDaisyWorld
    rng = np.random.RandomState(77)
    tiers = []
    for cfg in range(10):
        attrs, N, n, ramp, no_micro = draw(rng, sizes=(16, 64, 20))
        np.random.seed(100 + cfg)
        env = RLDaisyWorld(grid_dimension=N, n_agents=n, ramp_period=ramp)
        env.batch_size = 6
        for k, v in attrs.items():
            setattr(env, k, v)
        if no_micro:
            env.set_use_microclimate(False)
        env.reset()
        env.run(int(rng.randint(5, 120)), policy="greedy")
        st, _ = _check_f32_against_f64(env)
        tiers.append(st["fp64_tier"] / st["cells"])
    assert np.mean(tiers) < 0.25, tiers


@pytest.mark.parametrize("N,n", [(20, 3), (48, 4), (96, 2)])
def test_series_in_the_tile4_kernel_matches_oracle_per_step(N, n):
    """The 4x4-tile persistent kernel's in-kernel series mode (sides that are other multiples of 4), against the NumPy oracle
    started from the same reset state; more steps than one 16-step work item."""
    import json
    from therldaisyworld_b200 import RLDaisyWorld
    from oracle.daisy_numpy import OracleGreedy, env_from_golden
    np.random.seed(N)
    env = RLDaisyWorld(grid_dimension=N, n_agents=n)
    env.batch_size = 3
    env.reset()
    z = {"meta": json.dumps({"ctor": {"grid_dimension": N, "n_agents": n}, "attrs": {"batch_size": 3}}),
         "init_grid": env.grid.copy(), "init_agent_indices": env.agent_indices.copy(), "init_agent_states": env.agent_states.copy()}
    oenv, _ = env_from_golden(z)
    K = 37
    series = env.run_series(K, policy="greedy")
    agent = OracleGreedy()
    obs = oenv.get_obs(oenv.agent_indices)
    for t in range(K):
        obs, _, _, _ = oenv.step(agent(obs))
        np.testing.assert_allclose(series[t], [oenv.temp.mean(), oenv.grid[:, 1].mean(), oenv.grid[:, 2].mean()], rtol=1e-9)
    np.testing.assert_array_equal(env.grid, oenv.grid)
