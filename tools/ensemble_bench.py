"""Large ensembles on one GPU (BASELINE configs[2]/[3] shapes): B worlds of 64x64, device-side reset, whole lives on the device."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from therldaisyworld_b200 import RLDaisyWorld
from therldaisyworld_b200.ensemble import DeviceShard, simulate_lifespan

def run(B, policy, albedo=None, max_steps=100000, epsilon=None):
    np.random.seed(0)
    env = RLDaisyWorld(grid_dimension=64)
    env.batch_size = B
    if albedo:
        env.albedo_light, env.albedo_dark = albedo
    label = policy if epsilon is None else f"eps={epsilon}"
    t0 = time.perf_counter()
    env.reset_on_device(seed=13)
    if epsilon is not None:                       # Greedy(epsilon): one coin per step for the whole ensemble (greedy.py:23)
        env.set_epsilon(epsilon)
    env.synchronize()
    t1 = time.perf_counter()
    out = simulate_lifespan(DeviceShard(env), policy=policy, seed=7, device="cuda", max_steps=max_steps)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    steps = out["steps"]
    print(f"B={B} {label:10s} albedo={albedo}: reset {t1 - t0:.2f}s, {steps} steps in {t2 - t1:.2f}s -> "
          f"{B * 4096 * steps / (t2 - t1):.3e} cell-updates/s wall; biosphere {out['biosphere_lifespan_mean']:.2f}+-{out['biosphere_lifespan_sem']:.3f} "
          f"agents {out['agent_lifespan_mean']:.2f}+-{out['agent_lifespan_sem']:.3f}; mem {torch.cuda.mem_get_info()[0] / 2**30:.1f} GiB free", flush=True)

if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    for albedo in (None, (0.5, 0.5)):
        for policy in ("greedy", "antigreedy", "random"):
            run(B, policy, albedo)
        run(B, "eps_greedy", albedo, epsilon=0.5)     # SURVEY 8(d) cfg 3: the half-random condition
