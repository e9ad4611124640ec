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


def test_series_for_other_world_sizes_matches_oracle_per_step():
    """run_series on a shape outside the in-kernel series mode (16x16): same per-step means, sampled between one-step launches."""
    from oracle.daisy_numpy import OracleGreedy, env_from_golden
    z, meta = load_golden("greedy_n16_b4_todeath")
    env = product_env_from_golden(z, meta)
    K = 40
    series = env.run_series(K, policy="greedy")
    oenv, _ = env_from_golden(z)
    agent = OracleGreedy()
    obs = oenv.get_obs(oenv.agent_indices)
    for t in range(K):
        obs, _, _, _ = oenv.step(agent(obs))
        np.testing.assert_allclose(series[t], [oenv.temp.mean(), oenv.grid[:, 1].mean(), oenv.grid[:, 2].mean()], rtol=1e-9)
    np.testing.assert_array_equal(env.grid, oenv.grid)


def test_fp32_export_within_stated_tolerance():
    """fp32 mode of the fields: BASELINE.json's tolerance for fp32 is 1e-5 relative; the export is one rounding of the
    exact fp64 fields, so 6e-8 holds."""
    z, meta = load_golden("greedy_n64_b2_120")
    env = product_env_from_golden(z, meta)
    env.run(30, policy="greedy")
    g32, o32 = env.grid_f32(), env.observe_f32()
    assert g32.dtype == np.float32 and o32.dtype == np.float32
    np.testing.assert_allclose(g32, env.grid, rtol=1e-5, atol=0)
    np.testing.assert_allclose(g32, env.grid, rtol=6e-8, atol=0)
    np.testing.assert_allclose(o32, env.observe(), rtol=6e-8, atol=0)
