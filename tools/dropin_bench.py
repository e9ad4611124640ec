"""The drop-in step() path as a reference caller uses it: Greedy-like host policy, one env.step(action) per iteration
(materialising fp64 kernels, obs/reward/done downloaded every step) vs step_policy (action chosen on the device)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from therldaisyworld_b200 import RLDaisyWorld

def host_greedy(obs):
    food = (obs[..., 1, :, :] + obs[..., 2, :, :]).reshape(*obs.shape[:2], 9)[:, :, [3, 1, 7, 5]]
    return (4 + np.argmax(food, axis=-1))[..., None]

for B, N in ((32, 16), (1000, 8), (32, 64), (1000, 64)):
    np.random.seed(13)
    env = RLDaisyWorld(grid_dimension=N); env.batch_size = B; obs = env.reset()
    for mode in ("step(host greedy)", "step_policy(device greedy)"):
        for _ in range(3):
            obs, r, d, _i = env.step(host_greedy(obs)) if mode.startswith("step(") else env.step_policy("greedy")
        t0 = time.perf_counter()
        K = 50
        for _ in range(K):
            obs, r, d, _i = env.step(host_greedy(obs)) if mode.startswith("step(") else env.step_policy("greedy")
        dt = (time.perf_counter() - t0) / K
        print(f"N={N} B={B} {mode}: {dt * 1e3:.3f} ms/step -> {B / dt:.3e} env-steps/s, {B * N * N / dt:.3e} cell-updates/s", flush=True)
