"""Shared helpers for the parity tests."""
import json
import os

import numpy as np

from conftest import GOLDEN_DIR


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    return z, json.loads(str(z["meta"]))


def apply_attrs(env, attrs):
    for k, v in attrs.items():
        if k == "use_microclimate":
            env.set_use_microclimate(v)
        else:
            setattr(env, k, v)


def product_env_from_golden(z, meta, device=0):
    """Product (CUDA) environment put into the exact post-reset state recorded in a golden fixture."""
    from therldaisyworld_b200 import RLDaisyWorld
    state = np.random.get_state()
    env = RLDaisyWorld(device=device, **meta["ctor"])
    apply_attrs(env, meta["attrs"])
    env.reset()                                   # re-reads batch_size / n_agents / dim
    np.random.set_state(state)
    env.L = env.min_L
    env.dL = (env.max_L - env.min_L) / env.ramp_period
    env.step_count = 0
    env.grid = z["init_grid"].copy()
    env.agent_indices = z["init_agent_indices"].copy()
    env.agent_states = z["init_agent_states"].copy()
    return env


def golden_action(z, t):
    a = z["actions"][t]
    return None if (a.shape == (1, 1, 1) and a[0, 0, 0] == -1) else a


def uses_global_stream(z, meta):
    """Collision fixtures (N4): step() itself draws from the global np.random stream, interleaved with the policy's draws."""
    return "rng_key" in z.files and meta["ctor"].get("collision_mode", 0) == 1


def restore_stream(z):
    """Put the global legacy MT19937 stream where the reference had it right after reset()."""
    np.random.set_state(("MT19937", z["rng_key"], int(z["rng_pos"]), 0, 0.0))


def consume_policy_draws(z, meta, t):
    """Re-draw what the recorded policy drew from the global stream before step t (oracle/gen_golden.py::run_case)."""
    if meta["policy"]["kind"] == "randint":
        a = np.random.randint(9, size=(meta["B"], meta["n"], 1))
        np.testing.assert_array_equal(a, z["actions"][t])
    elif meta["policy"]["kind"] in ("greedy", "antigreedy"):
        np.random.rand()             # Greedy.__call__ flips its epsilon coin on every call (daisy/agents/greedy.py:23)
