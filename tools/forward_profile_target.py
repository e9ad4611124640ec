"""ncu target for the materialising single-step path (k_forward and friends): a few drop-in steps of a large batch."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from therldaisyworld_b200 import RLDaisyWorld
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
np.random.seed(13)
env = RLDaisyWorld(grid_dimension=64)
env.batch_size = B
env.reset()
for _ in range(6):
    env.step_policy("greedy")
