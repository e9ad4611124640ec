import sys, time, ctypes as C, numpy as np
sys.path.insert(0, '/root/repo')
from therldaisyworld_b200 import RLDaisyWorld
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
np.random.seed(13)
env = RLDaisyWorld(grid_dimension=64); env.batch_size = B; env.reset()
env.run(2, policy="greedy")           # off-lattice first step + warm
env.synchronize()
for rep in range(3):
    t = time.perf_counter(); env.run(384, policy="greedy"); env.synchronize(); dt = time.perf_counter() - t
    print(f"B={B} 384 steps: {dt*1e3:.2f} ms -> {B*4096*384/dt:.3e} cell-updates/s, {B*384/dt:.3e} env-steps/s")
c = C.c_uint64(); env._lib.dw_debug_slow_count(env._h, C.byref(c), 0); print("slow cells", c.value, c.value/(B*4096*384*3))
