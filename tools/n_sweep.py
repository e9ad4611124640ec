import sys, os, time, ctypes as C
sys.path.insert(0, os.getcwd())
import numpy as np
from therldaisyworld_b200 import RLDaisyWorld
from therldaisyworld_b200._lib import DwProfile
for N, B in ((8, 50000), (16, 20000), (32, 8000), (48, 4000), (64, 2000), (96, 1000), (128, 600), (150, 400), (33, 6000), (75, 1500), (100, 900), (110, 800)):
    np.random.seed(1)
    env = RLDaisyWorld(grid_dimension=N); env.batch_size = B; env.reset_on_device(seed=1); env.run(1, policy="greedy")
    env.synchronize(); t = time.perf_counter(); env.run(200, policy="greedy"); env.synchronize(); dt = time.perf_counter() - t
    print(f"N={N} B={B}: {B*N*N*200/dt:.3e} cell-updates/s", flush=True)
