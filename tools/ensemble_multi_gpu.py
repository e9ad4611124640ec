"""BASELINE configs[3]: a 1M-world random-agent ensemble (64x64, 4 agents) sharded over the GPUs of one node
(torchrun --nproc-per-node R tools/ensemble_multi_gpu.py [worlds_total]); whole lives on the device, the only traffic is
the per-segment 64-bit AND all-reduce of 'every world of my shard is done' masks and ONE all-reduce of the 8-double
lifespan statistics at the end (NCCL)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from therldaisyworld_b200 import RLDaisyWorld
from therldaisyworld_b200.ensemble import DeviceShard, shard_range, simulate_lifespan

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
total = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
lo, hi = shard_range(total, world, rank)
np.random.seed(0)
env = RLDaisyWorld(grid_dimension=64, device=local)
env.batch_size = hi - lo
for policy in ("random", "greedy"):
    env.reset_on_device(seed=13, world_offset=lo)
    env.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = simulate_lifespan(DeviceShard(env, world_offset=lo), policy=policy, seed=7, device="cuda")
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if rank == 0:
        print(f"ENSEMBLE worlds={out['worlds']} over {world} GPU(s) policy={policy}: {out['steps']} steps in {dt:.2f} s -> "
              f"{out['worlds'] * 4096 * out['steps'] / dt:.3e} cell-updates/s, {out['worlds'] * out['steps'] / dt:.3e} env-steps/s (wall, "
              f"incl. checkpoints/rewind); biosphere {out['biosphere_lifespan_mean']:.3f}+-{out['biosphere_lifespan_sem']:.4f} "
              f"agents {out['agent_lifespan_mean']:.3f}+-{out['agent_lifespan_sem']:.4f}", flush=True)
if world > 1:
    dist.destroy_process_group()
