"""ncu target: one ES population rollout (P members x 32 worlds of NxN) — launch list of the per-step sequence."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from therldaisyworld_b200.es import evaluate_population
P, N, steps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
members = np.random.RandomState(0).randn(P, 1808) * 0.5
np.random.seed(1)
evaluate_population(members, max_steps=steps, worlds_per_member=32, grid_dimension=N)
