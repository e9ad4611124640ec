"""GPU parity of the single-step (materialising) CUDA path, called through the C-ABI via the drop-in class.

Bar (BASELINE north_star): fp64 fields within 1e-9 relative of the reference; in practice -- and asserted
here -- value-identical (==) to the reference-recorded fixtures and to the C oracle, because the kernels
follow the oracle's IEEE operation order."""
import numpy as np
import pytest

from conftest import golden_names
from helpers import (golden_action, load_golden, product_env_from_golden, apply_attrs, consume_policy_draws, restore_stream,
                     uses_global_stream)

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", golden_names())
def test_step_replays_reference_trajectory(name):
    z, meta = load_golden(name)
    env = product_env_from_golden(z, meta)
    np.testing.assert_array_equal(env.get_obs(env.agent_indices), z["init_obs"])
    ck = {int(s): i for i, s in enumerate(z["ckpt_steps"])}
    stream = uses_global_stream(z, meta)          # collision_mode == 1: step() draws from the global np.random stream
    if stream:
        restore_stream(z)
    for t in range(meta["steps"]):
        assert env.L == z["L"][t]
        if stream:
            consume_policy_draws(z, meta, t)
        obs, reward, done, info = env.step(golden_action(z, t))
        assert info == {}
        np.testing.assert_array_equal(reward, z["reward"][t])
        np.testing.assert_array_equal(done, z["done"][t])
        assert reward.dtype == z["reward"][t].dtype and done.dtype == np.bool_
        if (t + 1) in ck or t % 37 == 0 or stream:
            np.testing.assert_array_equal(env.agent_indices, z["agent_indices"][t])
            np.testing.assert_array_equal(env.agent_states, z["agent_states"][t])
            np.testing.assert_array_equal(env.grid.sum(axis=(-2, -1)), z["chan_sum"][t])
        if (t + 1) in ck:
            np.testing.assert_array_equal(env.grid, z["ckpt_grid"][ck[t + 1]])
            np.testing.assert_array_equal(obs, z["ckpt_obs"][ck[t + 1]])
    assert env.L == z["L"][meta["steps"]]
    assert env.step_count == meta["final_step_count"]
    np.testing.assert_array_equal(env.agent_indices, z["agent_indices"][-1])
    np.testing.assert_array_equal(env.agent_states, z["agent_states"][-1])
    for key, attr in [("diag_temp", "temp"), ("diag_temp_light", "temp_light"), ("diag_temp_dark", "temp_dark"),
                      ("diag_temp_effective", "temp_effective"), ("diag_dead_temp", "dead_temp"),
                      ("diag_beta", "beta"), ("diag_beta_l", "beta_l"), ("diag_beta_d", "beta_d"),
                      ("diag_growth", "growth")]:
        got = getattr(env, attr)
        assert got.shape == z[key].shape, key
        np.testing.assert_allclose(got, z[key], rtol=1e-9, atol=1e-12, err_msg=key)   # FFT round-off in the reference


@pytest.mark.parametrize("name", ["cfg1_n16_b1_noagents_todeath", "greedy_n16_b4_todeath", "randint_n7_b3_n16_200",
                                  "greedy_n17_b2_params_200", "greedy_n64_b2_120"])
def test_reset_reproduces_reference_rng_order_and_init_fields(name):
    from therldaisyworld_b200 import RLDaisyWorld
    z, meta = load_golden(name)
    np.random.seed(meta["seed"])
    env = RLDaisyWorld(**meta["ctor"])
    apply_attrs(env, meta["attrs"])
    obs = env.reset()
    np.testing.assert_array_equal(env.agent_indices, z["init_agent_indices"])
    np.testing.assert_array_equal(env.agent_states, z["init_agent_states"])
    np.testing.assert_array_equal(env.grid[:, :3], z["init_grid"][:, :3])
    np.testing.assert_array_equal(env.grid[:, 6], 0.0)
    np.testing.assert_allclose(env.grid[:, 3:6], z["init_grid"][:, 3:6], rtol=1e-12)   # unrounded; FFT vs stencil
    np.testing.assert_allclose(obs, z["init_obs"], rtol=1e-12)
    assert env.step_count == 0 and env.L == env.min_L


def test_offlattice_states_match_c_oracle_bitwise():
    """Arbitrary covers, odd/small N (where the reference's FFT path fails), many agents, every action."""
    from therldaisyworld_b200 import RLDaisyWorld
    from oracle.daisy_c import COracleWorld
    rng = np.random.RandomState(7)
    for N, B, n in [(1, 2, 1), (2, 3, 2), (3, 2, 1), (4, 2, 3), (6, 3, 2), (9, 2, 5), (33, 2, 7), (64, 3, 4), (100, 1, 9)]:
        np.random.seed(N)
        env = RLDaisyWorld(grid_dimension=N, n_agents=n)
        env.batch_size = B
        env.reset()
        g = env.grid.copy()
        g[:, 1] = rng.rand(B, N, N) * 0.6
        g[:, 2] = rng.rand(B, N, N) * 0.4
        env.grid = g
        w = COracleWorld(env, grid=g)
        for t in range(6):
            action = rng.randint(9, size=(B, n, 1))
            o1, r1, d1, _ = env.step(action)
            o2, r2, d2, _ = w.step(action)
            np.testing.assert_array_equal(env.grid, w.grid)
            np.testing.assert_array_equal(o1, o2)
            np.testing.assert_array_equal(r1, r2)
            np.testing.assert_array_equal(d1, d2)
            np.testing.assert_array_equal(env.agent_indices, w.agent_indices)
            np.testing.assert_array_equal(env.agent_states[..., 0], w.agent_states.reshape(B, n))
        # unrounded diagnostics of the last forward are bit-identical too (same IEEE op order)
        diag = COracleWorld(env, grid=None)   # snapshot of current product state
        # recompute oracle diagnostics from the state BEFORE the last step is not available; instead check a
        # standalone forward on the current grid
        cur = env.grid.copy()
        out = env.forward(cur.copy())
        wd = COracleWorld(env, grid=cur)
        d = wd.forward_diag()
        np.testing.assert_array_equal(env.temp, d[:, 0:1])
        np.testing.assert_array_equal(env.temp_light, d[:, 1:2])
        np.testing.assert_array_equal(env.temp_dark, d[:, 2:3])
        np.testing.assert_array_equal(env.temp_effective, d[:, 3:4])
        np.testing.assert_array_equal(env.beta, d[:, 4:5])
        np.testing.assert_array_equal(env.beta_l, d[:, 5:6])
        np.testing.assert_array_equal(env.beta_d, d[:, 6:7])
        np.testing.assert_array_equal(env.growth, d[:, 7:9])
        wd.step(None) if n == 0 else None
        assert out.shape == cur.shape


def test_reference_smoke_tests_restated():
    """tests/daisy/test_daisy_world_rl.py:14-68 of the reference, against the drop-in."""
    from therldaisyworld_b200 import RLDaisyWorld
    env = RLDaisyWorld()
    a = env.grid
    b = env.forward(a)
    for ii in range(9):
        obs, reward, done, info = env.step(np.array([[[ii]]]))   # (1,1,1) action while B=32, n=4
    assert not done.mean()
    assert type(info) == dict
    assert 0.0 <= reward.mean()
    assert a.shape == b.shape
    assert obs.shape[1] == env.n_agents and obs.shape[0] == env.batch_size

    env = RLDaisyWorld()
    for c in (3, 4, 5):
        assert 0 < env.grid[:, c].mean()
    env.reset()
    for c in (3, 4, 5):
        assert 0 < env.grid[:, c].mean()
    obs, reward, done, info = env.step()
    for c in (3, 4, 5):
        assert 0 < env.grid[:, c].mean() and 0 < obs[:, :, c].mean()
    obs, reward, done, info = env.step(np.random.randint(9, size=(env.batch_size, env.n_agents, 1)))
    for c in (3, 4, 5):
        assert 0 < env.grid[:, c].mean() and 0 < obs[:, :, c].mean()


def test_host_mirror_round_trip_and_attribute_mutation():
    """Callers mutate attributes and arrays between steps; the device must see it."""
    from therldaisyworld_b200 import RLDaisyWorld
    from oracle.daisy_c import COracleWorld
    np.random.seed(3)
    env = RLDaisyWorld(grid_dimension=8, n_agents=2)
    env.batch_size = 3
    env.reset()
    g = env.grid
    g[:, 1:3] *= 0.5                      # in-place edit of the array we were handed
    env.agent_states[:, 0, 0] = 0.07      # nearly starved agent
    env.albedo_dark = 0.3
    env.dt = 0.5
    env.L = 1.1
    w = COracleWorld(env, grid=g)
    a = np.full((3, 2, 1), 6)
    o1, r1, d1, _ = env.step(a)
    o2, r2, d2, _ = w.step(a)
    np.testing.assert_array_equal(env.grid, w.grid)
    np.testing.assert_array_equal(o1, o2)
    np.testing.assert_array_equal(r1, r2)
    np.testing.assert_array_equal(d1, d2)
    assert env.L == w.L and env.step_count == w.step_count


def test_update_agents_and_get_obs_standalone():
    from therldaisyworld_b200 import RLDaisyWorld
    from oracle.daisy_numpy import OracleDaisyWorld
    np.random.seed(5)
    env = RLDaisyWorld(grid_dimension=9, n_agents=3)
    st = np.random.get_state()
    np.random.seed(5)
    ref = OracleDaisyWorld(grid_dimension=9, n_agents=3)
    np.random.set_state(st)
    np.testing.assert_array_equal(env.grid[:, :3], ref.grid[:, :3])
    ref.grid = env.grid.copy()
    action = np.array([[[5], [8], [2]]] * env.batch_size)
    env.update_agents(action)
    ref.update_agents(action)
    np.testing.assert_array_equal(env.grid, ref.grid)
    np.testing.assert_array_equal(env.agent_indices, ref.agent_indices)
    np.testing.assert_array_equal(env.agent_states, ref.agent_states)
    pos = np.random.randint(9, size=(5, 6, 2))
    np.testing.assert_array_equal(env.get_obs(pos), ref.get_obs(pos))


def test_collisions_match_numpy_oracle_on_crowded_worlds():
    """collision_mode == 1 beyond the recorded fixtures: crowded worlds (up to 60 agents on 4x4 .. 9x9, so cells with
    more than eight losers exercise NumPy's unrolled summation order), explicit actions, device policies and step(None);
    the drop-in and the NumPy oracle consume the same global stream and must agree exactly, stream position included."""
    from therldaisyworld_b200 import RLDaisyWorld
    from oracle.daisy_numpy import OracleDaisyWorld, OracleGreedy
    rng = np.random.RandomState(5)
    for N, B, n, penalty, mode in [(5, 4, 60, 0.5, "randint"), (9, 3, 25, 1.3, "greedy"), (7, 5, 9, 0.5, "none"),
                                   (8, 2, 33, 0.25, "partial")]:
        np.random.seed(100 + N)
        env = RLDaisyWorld(grid_dimension=N, n_agents=n, collision_mode=1)
        env.batch_size = B
        env.food_chain_penalty = penalty
        env.agent_gamma = 0.02
        env.reset()
        ref = OracleDaisyWorld(grid_dimension=N, n_agents=n, collision_mode=1)
        ref.batch_size, ref.food_chain_penalty, ref.agent_gamma = B, penalty, 0.02
        ref.reset()
        ref.grid, ref.agent_indices, ref.agent_states = env.grid.copy(), env.agent_indices.copy(), env.agent_states.copy()
        greedy = OracleGreedy(epsilon=0.0, greedy=True)
        obs_ref = ref.get_obs(ref.agent_indices)
        for t in range(40):
            if mode == "randint":
                action = rng.randint(9, size=(B, n, 1))
            elif mode == "partial":
                action = rng.randint(9, size=(B - 1, n - 3, 1))
            elif mode == "greedy":
                action = greedy(obs_ref)
            else:
                action = None
            state = np.random.get_state()
            if mode == "greedy" and t % 2:
                o1, r1, d1, _ = env.step_policy("greedy")          # decision on the device
            else:
                o1, r1, d1, _ = env.step(action)
            after_env = np.random.get_state()
            np.random.set_state(state)
            obs_ref, r2, d2, _ = ref.step(action)
            after_ref = np.random.get_state()
            assert after_env[2] == after_ref[2] and np.array_equal(after_env[1], after_ref[1])
            np.testing.assert_array_equal(env.agent_indices, ref.agent_indices)
            np.testing.assert_array_equal(env.agent_states, ref.agent_states)
            np.testing.assert_array_equal(o1, obs_ref)
            np.testing.assert_array_equal(r1, r2)
            np.testing.assert_array_equal(d1, d2)
        np.testing.assert_array_equal(env.grid, ref.grid)


def test_run_with_collisions_reproduces_reference_lifespans():
    """env.run() with collision_mode == 1: the loop is driven from the host (the noise is the caller's NumPy stream), the
    decisions, moves, collisions, forward and lifespan counters stay on the device. Against the trajectory recorded from the
    live reference with its Greedy policy. (The reference's Greedy also flips one epsilon coin per call from the same stream,
    which the device policy does not; in this fixture no collision is decided by the noise -- the oracle replay reproduces it
    with either alignment -- so the recorded states are the expected ones here as well.)"""
    z, meta = load_golden("collide_greedy_n8_b4_n10_150")
    env = product_env_from_golden(z, meta)
    env.reset_lifespans()
    restore_stream(z)
    steps, alive, hit = env.run(meta["steps"], policy="greedy")
    assert steps == meta["steps"]
    done_at, agents_done_at = env.lifespans()
    np.testing.assert_array_equal(done_at, z["done_at"])
    np.testing.assert_array_equal(agents_done_at, z["agents_done_at"])
    np.testing.assert_array_equal(env.agent_indices, z["agent_indices"][-1])
    np.testing.assert_array_equal(env.agent_states, z["agent_states"][-1])
    np.testing.assert_array_equal(env.grid, z["ckpt_grid"][-1])
    assert env.L == z["L"][meta["steps"]] and env.step_count == meta["final_step_count"]
    # the in-kernel series mode has no collision pass: run_series samples between host-driven steps instead
    out = env.run_series(3, policy="greedy")
    assert out.shape == (3, 3)
    np.testing.assert_allclose(out[-1], [env.temp.mean(), env.grid[:, 1].mean(), env.grid[:, 2].mean()], rtol=1e-12)


def test_collision_abi_state_machine_and_count_check():
    """Between dw_agents_begin and dw_agents_collide every other stepping call is refused (DW_E_STATE = -4), and a draw
    count that does not match the positions on the device fails the call (DW_E_INVALID = -1) instead of mis-assigning noise."""
    import ctypes as C
    from therldaisyworld_b200 import RLDaisyWorld
    np.random.seed(3)
    env = RLDaisyWorld(grid_dimension=4, n_agents=20, collision_mode=1)      # 20 agents on 16 cells: shared cells guaranteed
    env.batch_size = 2
    env.reset()
    env._push()
    lib, h = env._lib, env._h
    pos = np.empty((2, 20, 2), dtype=np.int64)
    assert lib.dw_agents_begin(h, None, 0, 0, -1, C.c_uint64(0), pos.ctypes.data_as(C.POINTER(C.c_int64))) == 0
    assert lib.dw_step_policy(h, 1, C.c_uint64(0)) == -4
    assert lib.dw_step(h, None, 0, 0) == -4
    assert lib.dw_agents_begin(h, None, 0, 0, -1, C.c_uint64(0), pos.ctypes.data_as(C.POINTER(C.c_int64))) == -4
    off = np.zeros(3, dtype=np.int32)                                         # claims "no shared cells"
    assert lib.dw_agents_collide(h, None, off.ctypes.data_as(C.POINTER(C.c_int32)), 0.5) == -1
    assert b"shared-cell counts" in lib.dw_last_error(h)
    # the failed call closed the pass; a regular step works again
    env._state_changed()
    obs, reward, done, _ = env.step(np.zeros((2, 20, 1), dtype=np.int64))
    assert obs.shape == (2, 20, 7, 3, 3)


def _window_obs(grid, pos, mask):
    """get_obs (daisy_world_rl.py:246-263) by plain indexing: 3x3 windows centred on pos, wrapped, masked."""
    B, m = pos.shape[:2]
    N = grid.shape[-1]
    out = np.zeros((B, m, 7, 3, 3))
    for b in range(B):
        for i in range(m):
            xs, ys = (pos[b, i, 0] + np.arange(-1, 2)) % N, (pos[b, i, 1] + np.arange(-1, 2)) % N
            out[b, i] = grid[b][:, xs][:, :, ys] * mask
    return out


def test_step_stays_lattice_resident_and_rebuilds_on_demand():
    """step() keeps the state on the packed lattice (no [B,7,N,N] materialisation); env.grid / diagnostics / get_obs are
    rebuilt lazily and change nothing; every value still equals the C oracle's."""
    from therldaisyworld_b200 import RLDaisyWorld
    from oracle.daisy_c import COracleWorld
    rng = np.random.RandomState(11)
    for N, B, n in [(64, 5, 4), (16, 7, 3), (20, 3, 5), (9, 2, 2)]:
        np.random.seed(N)
        env = RLDaisyWorld(grid_dimension=N, n_agents=n)
        env.batch_size = B
        obs0 = env.reset()
        assert env.residency() == dict(grid=False, lattice=False, cover_planes=True, pre=3), "reset() must stay lean"
        ref = COracleWorld(env)                                  # reads env.grid: materialises it once
        np.testing.assert_array_equal(obs0, ref.get_obs())
        for t in range(12):
            a = rng.randint(9, size=(B, n, 1))
            o1, r1, d1, _ = env.step(a)
            o2, r2, d2, _ = ref.step(a)
            res = env.residency()
            assert res["lattice"] and not res["grid"], f"step {t} materialised the grid: {res}"
            np.testing.assert_array_equal(o1, o2)
            np.testing.assert_array_equal(r1, r2)
            np.testing.assert_array_equal(d1, d2)
            if t in (0, 5, 11):
                np.testing.assert_array_equal(env.grid, ref.grid)
                assert env.residency()["lattice"], "reading env.grid must not drop the lattice"
            if t in (1, 6):                                      # get_obs at caller-supplied positions, still lean
                pos = rng.randint(N, size=(B, 2, 2))
                got = env.get_obs(pos)
                assert not env.residency()["grid"]
                np.testing.assert_array_equal(got, _window_obs(ref.grid, pos, env.neighborhood))
        o1, r1, d1, _ = env.step_policy("greedy", want_obs=False)
        ref.run(1, "greedy")
        assert o1 is None
        np.testing.assert_array_equal(r1[..., 0], np.clip(ref.agent_states.reshape(B, n), 0, None))
        np.testing.assert_array_equal(env.observe(), ref.get_obs())
        np.testing.assert_array_equal(env.grid, ref.grid)


def test_step_returns_fresh_arrays_from_the_pinned_pool():
    """The outputs come from page-locked blocks that are recycled only after the caller dropped them: arrays of earlier steps
    must never change, also when more than the pool's limit are kept alive."""
    from therldaisyworld_b200 import RLDaisyWorld
    np.random.seed(2)
    env = RLDaisyWorld(grid_dimension=16)
    env.batch_size = 6
    env.reset()
    kept, copies = [], []
    for t in range(40):
        out = env.step(np.random.randint(9, size=(6, 4, 1)))
        kept.append(out[:3])
        copies.append([x.copy() for x in out[:3]])
        if t % 3 == 0:
            kept.pop(0), copies.pop(0)                      # some are dropped: their blocks go back to the pool
    for got, want in zip(kept, copies):
        for g, w in zip(got, want):
            np.testing.assert_array_equal(g, w)
    assert out[2].dtype == np.bool_ and out[0].flags.writeable


def test_any_integer_action_is_read_like_the_reference():
    """daisy_world_rl.py:190-212 accepts any integer: a == 8 stays, a % 4 moves (Python modulo), a > 4 grazes."""
    from therldaisyworld_b200 import RLDaisyWorld
    from oracle.daisy_numpy import OracleDaisyWorld
    np.random.seed(8)
    env = RLDaisyWorld(grid_dimension=8, n_agents=6)
    env.batch_size = 4
    env.reset()
    ref = OracleDaisyWorld(grid_dimension=8, n_agents=6)
    ref.batch_size = 4
    ref.reset()
    ref.grid, ref.agent_indices, ref.agent_states = env.grid.copy(), env.agent_indices.copy(), env.agent_states.copy()
    rng = np.random.RandomState(1)
    for t in range(30):
        a = rng.randint(-9, 26, size=(4, 6, 1))
        o1, r1, d1, _ = env.step(a)
        o2, r2, d2, _ = ref.step(a)
        np.testing.assert_array_equal(env.agent_indices, ref.agent_indices)
        np.testing.assert_array_equal(env.agent_states, ref.agent_states)
        np.testing.assert_array_equal(o1, o2)
    np.testing.assert_array_equal(env.grid, ref.grid)
    acts = rng.randint(-9, 26, size=(10, 4, 6))
    env.run(10, policy="replay", actions=acts)
    for t in range(10):
        ref.step(acts[t][..., None])
    np.testing.assert_array_equal(env.grid, ref.grid)
    np.testing.assert_array_equal(env.agent_states, ref.agent_states)


def test_forward_leaves_the_live_state_alone():
    """env.forward(grid) is side-effect-free on the state (reference :434-461) also when the state is lattice-resident; its
    diagnostics are served until the state advances."""
    from therldaisyworld_b200 import RLDaisyWorld
    from oracle.daisy_c import COracleWorld
    np.random.seed(4)
    env = RLDaisyWorld(grid_dimension=16)
    env.batch_size = 3
    env.reset()
    ref = COracleWorld(env)
    env.run(9, policy="greedy")
    ref.run(9, "greedy")
    assert env.residency()["grid"] is False
    g = np.random.RandomState(0).rand(3, 7, 16, 16) * 0.3
    out = env.forward(g.copy())
    wd = COracleWorld(env, grid=g.copy(), agent_indices=ref.agent_indices, agent_states=ref.agent_states)
    np.testing.assert_array_equal(env.temp, wd.forward_diag()[:, 0:1])
    np.testing.assert_array_equal(env.grid, ref.grid)                      # still there, still right
    a = np.full((3, 4, 1), 7)
    o1, _, _, _ = env.step(a)
    o2, _, _, _ = ref.step(a)
    np.testing.assert_array_equal(o1, o2)
    np.testing.assert_array_equal(env.grid, ref.grid)
    assert out.shape == g.shape
