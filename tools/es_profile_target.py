"""One ES generation (P members x 32 worlds of NxN, 768 max steps) for the launch list: python tools/es_profile_target.py [P] [N]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from therldaisyworld_b200.es import evaluate_population
P = int(sys.argv[1]) if len(sys.argv) > 1 else 64
N = int(sys.argv[2]) if len(sys.argv) > 2 else 16
members = np.random.RandomState(0).randn(P, 1808) * 0.5
members[0] = 0.0
env = None
for rep in range(2):
    np.random.seed(1)
    t0 = time.perf_counter()
    fitness, total_steps, member_steps, env = evaluate_population(members, max_steps=768, worlds_per_member=32, env=env, grid_dimension=N)
    t1 = time.perf_counter()
    f2, _, ms2, env = evaluate_population(members, max_steps=768, worlds_per_member=32, env=env, grid_dimension=N, device_reset_seed=3)
    t2 = time.perf_counter()
    print(f"rep {rep}: host-draw generation {1e3 * (t1 - t0):.1f} ms ({int(member_steps.max())} steps), device-draw generation {1e3 * (t2 - t1):.1f} ms "
          f"({int(ms2.max())} steps)", flush=True)
