"""Harness shared by tests/test_gpu_reference_callers.py (CUDA drop-in under the reference's own callers) and
tests/test_reference_harness.py (CPU self-check of the harness with the reference's own class): imports the UNMODIFIED
reference from the git-ignored baseline/_ref (baseline/install_ref.py), optionally aliasing its environment class."""
import importlib
import os
import sys
import tempfile
import warnings

import numpy as np
import pytest

from conftest import ROOT
from helpers import apply_attrs, load_golden

REF = os.path.join(ROOT, "baseline", "_ref")


def have_ref():
    return os.path.exists(os.path.join(REF, "daisy", "nn", "functional.py"))


def import_reference(alias=None):
    """Generator: imports the reference from baseline/_ref (mpi4py stubbed); with `alias`, daisy.daisy_world_rl.RLDaisyWorld is
    replaced by it BEFORE the callers are imported (the one-line switch of INTEGRATION.md section 1). Cleans up afterwards."""
    if not have_ref():
        pytest.skip("baseline/_ref not installed (python baseline/install_ref.py needs /root/reference)")
    stub = tempfile.mkdtemp()                      # mpi4py is absent: sges/cmaes only touch MPI.COMM_WORLD at import time
    os.makedirs(os.path.join(stub, "mpi4py"))
    with open(os.path.join(stub, "mpi4py", "__init__.py"), "w") as f:
        f.write("from . import MPI\n")
    with open(os.path.join(stub, "mpi4py", "MPI.py"), "w") as f:
        f.write("class _Comm:\n    def Get_rank(self): return 0\n    def Get_size(self): return 1\nCOMM_WORLD = _Comm()\n")
    saved_path = list(sys.path)
    sys.path.insert(0, stub)
    sys.path.insert(0, REF)
    warnings.filterwarnings("ignore", category=DeprecationWarning)
    import daisy.daisy_world_rl as ref_env_mod
    assert os.path.realpath(ref_env_mod.__file__).startswith(os.path.realpath(REF))
    ref_class = ref_env_mod.RLDaisyWorld
    env_class = ref_class if alias is None else alias
    ref_env_mod.RLDaisyWorld = env_class
    mods = {}
    for name in ("daisy.agents.greedy", "daisy.agents.mlp", "daisy.evo.sges", "daisy.evo.cmaes"):
        m = importlib.import_module(name)
        assert m.RLDaisyWorld is env_class          # `from daisy.daisy_world_rl import RLDaisyWorld` saw the alias
        mods[name] = m
    yield dict(env=env_class, ref_class=ref_class, greedy=mods["daisy.agents.greedy"].Greedy, mlp=mods["daisy.agents.mlp"].MLP,
               sges=mods["daisy.evo.sges"].SimpleGaussianES)
    ref_env_mod.RLDaisyWorld = ref_class
    sys.path[:] = saved_path
    for k in [k for k in sys.modules if k == "daisy" or k.startswith("daisy.") or k.startswith("mpi4py")]:
        del sys.modules[k]


def drive(ref, name, agent, steps=None):
    """oracle/gen_golden.py::run_case with ref["env"] as the environment class: every action the reference policy picks and
    everything step() returns must equal the all-reference recording."""
    z, meta = load_golden(name)
    np.random.seed(meta["seed"])
    env = ref["env"](**meta["ctor"])
    apply_attrs(env, meta["attrs"])
    obs = env.reset()
    np.testing.assert_array_equal(env.agent_indices, z["init_agent_indices"])
    np.testing.assert_allclose(obs, z["init_obs"], rtol=1e-12)          # unrounded initial temperatures: FFT vs stencil
    ck = {int(s): i for i, s in enumerate(z["ckpt_steps"])}
    T = meta["steps"] if steps is None else min(steps, meta["steps"])
    for t in range(T):
        action = agent(obs)
        np.testing.assert_array_equal(np.asarray(action), z["actions"][t], err_msg=f"{name}: action of step {t}")
        obs, reward, done, info = env.step(action)
        np.testing.assert_array_equal(reward, z["reward"][t])
        np.testing.assert_array_equal(done, z["done"][t])
        if (t + 1) in ck:
            np.testing.assert_array_equal(env.agent_indices, z["agent_indices"][t])
            np.testing.assert_array_equal(env.agent_states, z["agent_states"][t])
            np.testing.assert_array_equal(env.grid, z["ckpt_grid"][ck[t + 1]])
            np.testing.assert_array_equal(obs, z["ckpt_obs"][ck[t + 1]])
    np.testing.assert_array_equal(env.agent_indices, z["agent_indices"][T - 1])
    np.testing.assert_array_equal(env.agent_states, z["agent_states"][T - 1])
    assert env.L == z["L"][T]
    return env


def greedy_agent(ref, kind):
    agent = ref["greedy"]()
    agent.greedy = kind != "antigreedy"
    agent.epsilon = {"greedy": 0.0, "antigreedy": 0.0, "random": 1.0, "half_random": 0.5}[kind]
    return agent
