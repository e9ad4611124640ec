#!/usr/bin/env python3
"""Golden fixture for next-row N2 (ES fitness rollout) from the UNMODIFIED reference (test infrastructure only).

Runs daisy.evo.sges.SimpleGaussianES.get_fitness (daisy/evo/sges.py:144-181) of the live reference for a small
population with fixed MLP weights.  mpi4py is absent in this container; the harness only touches MPI.COMM_WORLD at import
time when num_workers == 0, so a two-line stub module stands in for it.

Usage: python oracle/gen_golden_es.py [--ref /root/reference] [--out tests/golden]"""
import argparse
import json
import os
import sys
import tempfile
import warnings

import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(os.path.dirname(__file__), "..", "tests", "golden"))
    args = ap.parse_args()
    stub = tempfile.mkdtemp()
    os.makedirs(os.path.join(stub, "mpi4py"))
    open(os.path.join(stub, "mpi4py", "__init__.py"), "w").write("from . import MPI\n")
    open(os.path.join(stub, "mpi4py", "MPI.py"), "w").write(
        "class _Comm:\n    def Get_rank(self): return 0\n    def Get_size(self): return 1\nCOMM_WORLD = _Comm()\n")
    sys.path.insert(0, stub)
    sys.path.insert(0, args.ref)
    warnings.filterwarnings("ignore")
    from daisy.evo.sges import SimpleGaussianES

    P, max_steps, dim = 4, 48, 16
    np.random.seed(31)
    es = SimpleGaussianES(population_size=P, max_steps=max_steps, grid_dimension=dim)
    params = np.zeros((P, 1808))
    import glob
    base = np.array(json.load(open(glob.glob(os.path.join(args.ref, "results", "cmaes_exp_002", "*best_agent_gen127.json"))[0]))["parameters"])
    params[2] = base + np.random.RandomState(2).randn(1808) * 2.0    # grazes (actions 5/6); members 0, 1: zero weights ->
    params[3] = base                                                 # action 0 forever -> starve (early all_done for member 1)
    for k, m in enumerate(es.population):
        m.set_parameters(params[k].copy())
    np.random.seed(77)                                            # the resets inside get_fitness draw from here, in order
    fitness, total_steps, done_at, steps_run = [], [], [], []
    for i in range(P):
        f, ts, da = es.get_fitness(agent_idx=i, adversary_idx=0)
        fitness.append(f); total_steps.append(np.asarray(ts)); done_at.append(np.asarray(da)); steps_run.append(es.env.step_count)
    meta = dict(P=P, max_steps=max_steps, grid_dimension=dim, batch_size=int(es.env.batch_size), n_agents=int(es.env.n_agents),
                reset_seed=77, adversary_idx=0, steps_run=[int(s) for s in steps_run], numpy=np.__version__)
    path = os.path.join(args.out, "es_fitness_p4_n16.npz")
    np.savez_compressed(path, params=params, fitness=np.array(fitness), total_steps=np.stack(total_steps), done_at=np.stack(done_at),
                        meta=np.array(json.dumps(meta)))
    print(path, "fitness", fitness, "steps_run", steps_run)


if __name__ == "__main__":
    main()
