"""ES fitness rollout of a whole population on one GPU (next-row N2): P members x 32 worlds, default 16x16 grid, 768 steps."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from therldaisyworld_b200.es import evaluate_population

for P, N in ((16, 16), (64, 16), (256, 16), (64, 64)):
    members = np.random.RandomState(0).randn(P, 1808) * 0.5
    members[0] = 0.0
    env = None
    for rep in range(2):
        np.random.seed(1)
        t0 = time.perf_counter()
        fitness, total_steps, member_steps, env = evaluate_population(members, max_steps=768, worlds_per_member=32, env=env, grid_dimension=N)
        dt = time.perf_counter() - t0
    for rep in range(2):
        t0 = time.perf_counter()
        f2, _, ms2, env = evaluate_population(members, max_steps=768, worlds_per_member=32, env=env, grid_dimension=N, device_reset_seed=3)
        dt_dev = time.perf_counter() - t0
    print(f"P={P} N={N}: device-side reset draws: {int(ms2.max())} steps, {dt_dev * 1e3:.1f} ms wall")
    steps = int(member_steps.max())
    print(f"P={P} N={N}: {steps} steps, {dt * 1e3:.1f} ms wall (incl. host reset draws + upload) -> {P * 32 * steps / dt:.3e} env-steps/s, "
          f"{P * 32 * N * N * steps / dt:.3e} cell-updates/s; fitness[:3]={fitness[:3]}", flush=True)
