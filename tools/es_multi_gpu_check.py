"""torchrun --nproc-per-node R tools/es_multi_gpu_check.py: the population fitness rollout sharded over R GPUs reproduces
the values recorded from the live reference's sequential SimpleGaussianES.get_fitness loop (tests/golden/es_fitness_p4_n16.npz)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
from therldaisyworld_b200.es import evaluate_population_sharded

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
z = np.load(os.path.join(ROOT, "tests", "golden", "es_fitness_p4_n16.npz"))
meta = json.loads(str(z["meta"]))
np.random.seed(meta["reset_seed"])
fit = evaluate_population_sharded(z["params"], adversary_idx=meta["adversary_idx"], max_steps=meta["max_steps"],
                                  worlds_per_member=meta["batch_size"], device=local, grid_dimension=meta["grid_dimension"])
if rank == 0:
    ok = np.allclose(fit, z["fitness"], rtol=1e-12, atol=0)
    print(f"ES SHARDED over {world} GPUs: {'OK' if ok else 'MISMATCH'} {fit} vs {z['fitness']}", flush=True)
dist.destroy_process_group()
