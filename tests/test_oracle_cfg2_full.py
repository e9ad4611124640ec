"""BASELINE config 2 at FULL size (1000 worlds, 64x64, 4 greedy agents, seed 13, run to all-dead = 470 steps,
1.9e9 cell-updates): the C oracle reproduces the lifespans and final-state checksums recorded from the live
reference (oracle/gen_golden_cfg2.py, ~20 min of reference time).  ~1 min on 8 cores."""
import numpy as np

from helpers import load_golden


def cfg2_initial_oracle_env():
    from oracle.daisy_numpy import OracleDaisyWorld
    np.random.seed(13)
    env = OracleDaisyWorld(grid_dimension=64)
    env.batch_size = 1000
    env.reset()
    return env


def test_c_oracle_reproduces_reference_cfg2_lifespans():
    from oracle.daisy_c import COracleWorld
    z, meta = load_golden("big_cfg2_greedy_n64_b1000")
    env = cfg2_initial_oracle_env()
    np.testing.assert_array_equal(env.grid[:, 1:3].sum(axis=(-2, -1)), z["init_daisy_sum"])
    np.testing.assert_array_equal(env.agent_indices, z["init_agent_indices"])
    w = COracleWorld(env)
    steps, done_at, agents_done_at = w.run(100000, "greedy", stop_all_done=True)
    assert steps == meta["steps"] == 470
    np.testing.assert_array_equal(done_at, z["done_at"])
    np.testing.assert_array_equal(agents_done_at, z["agents_done_at"])
    np.testing.assert_array_equal(w.grid.sum(axis=(-2, -1)), z["final_chan_sum"])
    np.testing.assert_array_equal(w.agent_states.reshape(1000, 4, 1), z["final_agent_states"])
    np.testing.assert_array_equal(w.agent_indices, z["final_agent_indices"])
