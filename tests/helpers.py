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
