"""The reference's OWN callers, unmodified, driven through the CUDA drop-in.

``baseline/_ref`` holds the unmodified reference (``baseline/install_ref.py``; git-ignored, travels to the GPU box).
``daisy.daisy_world_rl.RLDaisyWorld`` is aliased to ``therldaisyworld_b200.RLDaisyWorld`` -- the one-line switch a
maintainer makes -- and then the reference's ``Greedy`` (daisy/agents/greedy.py:5-36), ``MLP`` (daisy/agents/mlp.py:97-116),
``SimpleGaussianES.get_fitness`` (daisy/evo/sges.py:144-181) and its unit tests (tests/daisy/test_daisy_world_rl.py:14-68,
tests/daisy/agents/test_greedy.py) run as they are: ``env.step(action)`` once per step, NumPy in and out, the global
``np.random`` stream shared between env and policy.  Every action the reference policy picks, every reward / done / position
and the checkpointed grids must equal what the all-reference run recorded (tests/golden, oracle/gen_golden*.py).
Skipped when ``baseline/_ref`` is absent."""
import json
import os
import unittest

import numpy as np
import pytest

from conftest import GOLDEN_DIR
from helpers import load_golden
from ref_harness import REF, drive, greedy_agent, import_reference

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref():
    import therldaisyworld_b200 as fast
    yield from import_reference(alias=fast.RLDaisyWorld)


@pytest.mark.parametrize("name,kind", [("greedy_n16_b4_todeath", "greedy"), ("antigreedy_n8_b16_todeath", "antigreedy"),
                                       ("halfrandom_n16_b4_300", "half_random"), ("random_n8_b8_todeath", "random"),
                                       ("greedy_n64_b2_120", "greedy"), ("greedy_moore_n64_b2_n6_60", "greedy"),
                                       ("rampupdown_n8_b2_100", "greedy")])
def test_unmodified_greedy_drives_the_cuda_environment(ref, name, kind):
    drive(ref, name, greedy_agent(ref, kind))


@pytest.mark.parametrize("name", ["mlp_n16_b4_mixed_150", "mlp_n64_b2_n6_60", "mlp_moore_n16_b3_100", "mlp_circular_n16_b2_60"])
def test_unmodified_mlp_drives_the_cuda_environment(ref, name):
    z, _ = load_golden(name)
    agent = ref["mlp"]()
    agent.set_parameters(z["mlp_params"])
    drive(ref, name, agent)


def test_unmodified_collision_mode_loop(ref):
    """collision_mode == 1 draws its noise from the global stream inside step(): the unmodified Greedy (one coin per call) and
    the CUDA environment must interleave their draws exactly like the all-reference run."""
    drive(ref, "collide_greedy_n8_b4_n10_150", ref["greedy"]())


def test_unmodified_sges_get_fitness_on_the_cuda_environment(ref):
    """SimpleGaussianES builds its environment from the aliased class (sges.py:27-29) and steps it once per step (:170)."""
    z = np.load(os.path.join(GOLDEN_DIR, "es_fitness_p4_n16.npz"))
    meta = json.loads(str(z["meta"]))
    np.random.seed(31)
    es = ref["sges"](population_size=meta["P"], max_steps=meta["max_steps"], grid_dimension=meta["grid_dimension"])
    assert isinstance(es.env, ref["env"])
    for k, m in enumerate(es.population):
        m.set_parameters(z["params"][k].copy())
    np.random.seed(meta["reset_seed"])
    for i in range(meta["P"]):
        f, ts, da = es.get_fitness(agent_idx=i, adversary_idx=meta["adversary_idx"])
        assert f == z["fitness"][i]                       # same rewards, same NumPy summation: identical, not just close
        np.testing.assert_array_equal(np.asarray(ts), z["total_steps"][i])
        np.testing.assert_array_equal(np.asarray(da), z["done_at"][i])
        assert es.env.step_count == meta["steps_run"][i]


@pytest.mark.parametrize("path,case", [("daisy/test_daisy_world_rl.py", "TestRLDaisyWorld"), ("daisy/agents/test_greedy.py", "TestGreedy")])
def test_reference_unit_tests_pass_on_the_cuda_environment(ref, path, case):
    """The reference's own unit tests (tests/daisy/...), loaded from baseline/_ref/ref_tests and run unchanged."""
    import importlib.util
    file = os.path.join(REF, "ref_tests", path)
    spec = importlib.util.spec_from_file_location("ref_" + os.path.basename(path)[:-3], file)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    if hasattr(mod, "RLDaisyWorld"):
        assert mod.RLDaisyWorld is ref["env"]
    suite = unittest.defaultTestLoader.loadTestsFromTestCase(getattr(mod, case))
    assert suite.countTestCases() > 0
    np.random.seed(0)
    result = unittest.TextTestRunner(verbosity=0).run(suite)
    assert result.wasSuccessful(), result.failures + result.errors
