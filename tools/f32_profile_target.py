"""ncu target for the fp32 mode: a lattice-resident ensemble, then env.grid_f32() (k_forward_f32 + k_stamp_f32): python tools/f32_profile_target.py [B]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from therldaisyworld_b200 import RLDaisyWorld
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
np.random.seed(13)
env = RLDaisyWorld(grid_dimension=64)
env.batch_size = B
env.reset_on_device(seed=13)
env.run(40, policy="greedy")
for _ in range(3):
    g = env.grid_f32()
print(g.shape, g.dtype, env.f32_stats())
