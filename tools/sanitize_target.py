"""Quick end-to-end target (also usable under a sanitizer where one is available): a few steps through every fused kernel (sub-64, 64x64 persistent / simple, multiple-of-4 tile,
one-cell-per-thread), the materialising step, the collision pass and the MLP policy, at sizes a sanitizer run finishes quickly."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from therldaisyworld_b200 import RLDaisyWorld

for N, B, n in ((8, 70, 4), (16, 19, 9), (32, 5, 33), (64, 6, 4), (64, 2, 40), (20, 3, 3), (96, 2, 5), (33, 2, 7)):
    np.random.seed(N)
    env = RLDaisyWorld(grid_dimension=N, n_agents=n)
    env.batch_size = B
    env.reset()
    env.reset_lifespans()
    env.run(12, policy="greedy")
    env.run(5, policy="random", seed=3)
    env.step_policy("antigreedy")
    g = env.grid
    print(N, B, n, "ok", float(g[:, 1:3].mean()), flush=True)
np.random.seed(1)
env = RLDaisyWorld(grid_dimension=6, n_agents=30, collision_mode=1)
env.batch_size = 3
env.reset()
for t in range(5):
    env.step(np.random.randint(9, size=(3, 30, 1)))
print("collide ok", flush=True)
env = RLDaisyWorld(grid_dimension=16)
env.batch_size = 4
env.reset()
env.set_mlp(np.random.RandomState(0).randn(1808))
env.run(6, policy="mlp")
print("mlp ok", float(env.agent_states.mean()), flush=True)
if len(sys.argv) > 1 and sys.argv[1] == "series":
    env = RLDaisyWorld(grid_dimension=64)
    env.batch_size = 3
    env.reset()
    print(env.run_series(6)[-1])
