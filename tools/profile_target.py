"""Small, deterministic invocation of the hot path for ncu: 1000 worlds, 64x64, greedy, 1 + 2*64 steps."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from therldaisyworld_b200 import RLDaisyWorld
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 129
np.random.seed(13)
env = RLDaisyWorld(grid_dimension=64); env.batch_size = B; env.reset()
env.reset_lifespans()
print(env.run(K, policy="greedy"))
print(env.lifespans()[0][:8])
