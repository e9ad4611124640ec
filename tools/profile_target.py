"""Small, deterministic invocation of the hot path for ncu: B worlds (default 1000) of NxN (default 64), greedy, K steps
(default 1 + 2*64): python tools/profile_target.py [B] [K] [N]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from therldaisyworld_b200 import RLDaisyWorld
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 129
N = int(sys.argv[3]) if len(sys.argv) > 3 else 64
np.random.seed(13)
env = RLDaisyWorld(grid_dimension=N); env.batch_size = B; env.reset()
env.reset_lifespans()
print(env.run(K, policy="greedy"))
print(env.lifespans()[0][:8])
