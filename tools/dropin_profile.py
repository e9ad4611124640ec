"""Where the time of one lattice-resident env.step() goes (N = 64, B = 1000): Python + C call overhead (B = 1 as the floor),
the K = 1 fused launch, the observation kernel, and the device->host copy of the packed outputs (pinned vs pageable)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from therldaisyworld_b200 import RLDaisyWorld


def t_of(fn, iters=300, warm=30):
    for _ in range(warm):
        fn()
    t0 = time.perf_counter()
    for _ in range(iters):
        fn()
    return (time.perf_counter() - t0) / iters * 1e6


for B in (1, 1000):
    np.random.seed(13)
    env = RLDaisyWorld(grid_dimension=64); env.batch_size = B; env.reset()
    env.step_policy("greedy")
    print(f"B={B}: step_policy no_obs {t_of(lambda: env.step_policy('greedy', want_obs=False)):.1f} us, with obs "
          f"{t_of(lambda: env.step_policy('greedy')):.1f} us, observe() alone (pageable) {t_of(lambda: env.observe()):.1f} us, "
          f"_push alone {t_of(lambda: env._push()):.2f} us, _out_views alone {t_of(lambda: env._out_views(True)):.2f} us", flush=True)
    if len(sys.argv) > 1 and sys.argv[1] == "once":
        break
for mb in (0.036, 0.5, 2.0, 8.0):
    n = int(mb * 1e6)
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    hp = torch.empty(n, dtype=torch.uint8).pin_memory()
    hq = torch.empty(n, dtype=torch.uint8)
    def cp(dst):
        dst.copy_(d, non_blocking=True); torch.cuda.synchronize()
    print(f"D2H {mb} MB: pinned {t_of(lambda: cp(hp)):.1f} us, pageable {t_of(lambda: cp(hq)):.1f} us", flush=True)

# the C call alone (no Python array work): packed step into one pre-allocated pinned block / one pageable block
import ctypes as C
from therldaisyworld_b200._lib import DwClock
np.random.seed(13)
env = RLDaisyWorld(grid_dimension=64); env.batch_size = 1000; env.reset(); env.step_policy("greedy")
lib, h = env._lib, env._h
lay = (C.c_int64 * 4)(); lib.dw_step_out_layout(h, lay)
pin = C.c_void_p(); lib.dw_host_alloc(C.c_uint64(lay[3]), C.byref(pin))
page = np.empty(lay[3], dtype=np.uint8)
clk = DwClock()
for name, ptr in (("pinned", pin), ("pageable", page.ctypes.data_as(C.c_void_p))):
    for want in (0, 1):
        f = lambda: lib.dw_step_packed(h, None, 0, 0, 1, C.c_uint64(0), want, ptr, C.byref(clk))
        print(f"C call dw_step_packed B=1000 {name} want_obs={want}: {t_of(f):.1f} us", flush=True)
print("pool live/free:", env._pool._live, {k: len(v) for k, v in env._pool._free.items()})
