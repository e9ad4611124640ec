"""ncu target for the in-kernel MLP policy: python tools/mlp_profile_target.py [B] [K]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from therldaisyworld_b200 import RLDaisyWorld
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 129
np.random.seed(13)
env = RLDaisyWorld(grid_dimension=64)
env.batch_size = B
env.reset()
env.set_mlp(np.random.RandomState(0).randn(1808))
env.run(2, policy="mlp")
print(env.run(K, policy="mlp"))
