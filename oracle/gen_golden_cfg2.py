#!/usr/bin/env python3
"""BASELINE config 2 at full size from the live reference: seed 13, RLDaisyWorld(grid_dimension=64),
batch_size=1000, Greedy(eps=0), notebook lifespan loop to all-dead.  The initial state is NOT stored
(65 MB): it is regenerated from the seed by the oracle's reset(), whose RNG order is pinned by
tests/test_oracle_golden.py.  Stored: lifespans, per-world checksums of the final grid, steps, and the
initial l/d channel sums for a self-check.   ~20 min of reference time.
Usage: python oracle/gen_golden_cfg2.py [B] [policy]"""
import json, os, sys, time, warnings
import numpy as np
sys.path.insert(0, "/root/reference")
warnings.filterwarnings("ignore")
from daisy.daisy_world_rl import RLDaisyWorld
from daisy.agents.greedy import Greedy

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
policy = sys.argv[2] if len(sys.argv) > 2 else "greedy"
seed = 13
np.random.seed(seed)
env = RLDaisyWorld(grid_dimension=64)
env.batch_size = B
obs = env.reset()
agent = Greedy(); agent.greedy = policy != "antigreedy"
init_sum = env.grid[:, 1:3].sum(axis=(-2, -1))
init_agents = env.agent_indices.copy()
done_at = np.zeros((B,), dtype=np.int64); agents_done_at = np.zeros((B, 4, 1), dtype=np.int64)
t0 = time.time(); steps = 0
while True:
    action = agent(obs)
    obs, reward, done, info = env.step(action)
    steps += 1
    grid_done = env.grid[:, 1:3].max(axis=(1, 2, 3)) <= 0.005
    done_at += 1 - 1 * grid_done
    agents_done_at += 1 - 1 * done
    if grid_done.mean() == 1.0:
        break
    if steps % 50 == 0:
        print(steps, time.time() - t0, flush=True)
meta = dict(seed=seed, B=B, N=64, n=4, policy=policy, steps=steps, seconds=time.time() - t0, numpy=np.__version__)
out = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", f"big_cfg2_{policy}_n64_b{B}.npz")
os.makedirs(os.path.dirname(out), exist_ok=True)
np.savez_compressed(out, meta=np.array(json.dumps(meta)), done_at=done_at, agents_done_at=agents_done_at,
                    final_chan_sum=env.grid.sum(axis=(-2, -1)), final_agent_states=env.agent_states,
                    final_agent_indices=env.agent_indices, init_daisy_sum=init_sum, init_agent_indices=init_agents)
print("wrote", out, meta)
