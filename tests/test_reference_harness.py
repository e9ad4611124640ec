"""CPU self-check of tests/ref_harness.py: the SAME harness that tests/test_gpu_reference_callers.py points at the CUDA class,
pointed at the reference's own class, reproduces the committed recordings -- so a GPU-side mismatch is the product's, not the
harness's. Uses the unmodified reference in baseline/_ref (skipped when absent)."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR
from helpers import load_golden
from ref_harness import drive, greedy_agent, import_reference


@pytest.fixture(scope="module")
def ref():
    yield from import_reference(alias=None)


@pytest.mark.parametrize("name,kind", [("halfrandom_n16_b4_300", "half_random"), ("collide_greedy_n8_b4_n10_150", "greedy")])
def test_harness_with_reference_class_reproduces_the_recording(ref, name, kind):
    drive(ref, name, greedy_agent(ref, kind), steps=120)


def test_harness_mlp_with_reference_class(ref):
    z, _ = load_golden("mlp_moore_n16_b3_100")
    agent = ref["mlp"]()
    agent.set_parameters(z["mlp_params"])
    drive(ref, "mlp_moore_n16_b3_100", agent, steps=60)


def test_harness_sges_with_reference_class(ref):
    z = np.load(os.path.join(GOLDEN_DIR, "es_fitness_p4_n16.npz"))
    meta = json.loads(str(z["meta"]))
    np.random.seed(31)
    es = ref["sges"](population_size=meta["P"], max_steps=meta["max_steps"], grid_dimension=meta["grid_dimension"])
    for k, m in enumerate(es.population):
        m.set_parameters(z["params"][k].copy())
    np.random.seed(meta["reset_seed"])
    for i in range(meta["P"]):
        f, ts, da = es.get_fitness(agent_idx=i, adversary_idx=meta["adversary_idx"])
        assert f == z["fitness"][i]
        np.testing.assert_array_equal(np.asarray(ts), z["total_steps"][i])
        assert es.env.step_count == meta["steps_run"][i]
