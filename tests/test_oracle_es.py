"""Next-row N2: the oracle's restatement of SimpleGaussianES.get_fitness (daisy/evo/sges.py:144-181) reproduces the
fitness values, step totals and stopping steps recorded from the live reference (oracle/gen_golden_es.py)."""
import json
import os

import numpy as np

from conftest import GOLDEN_DIR
from oracle.daisy_numpy import OracleDaisyWorld, OracleMLP, es_get_fitness


def test_oracle_es_fitness_matches_reference():
    z = np.load(os.path.join(GOLDEN_DIR, "es_fitness_p4_n16.npz"))
    meta = json.loads(str(z["meta"]))
    env = OracleDaisyWorld(grid_dimension=meta["grid_dimension"])
    pop = [OracleMLP(p) for p in z["params"]]
    np.random.seed(meta["reset_seed"])
    for i in range(meta["P"]):
        f, ts, da, steps = es_get_fitness(env, pop[i], pop[meta["adversary_idx"]], max_steps=meta["max_steps"])
        assert f == z["fitness"][i]
        np.testing.assert_array_equal(ts, z["total_steps"][i])
        np.testing.assert_array_equal(da, z["done_at"][i])
        assert steps == meta["steps_run"][i]
