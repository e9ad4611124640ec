"""torchrun --nproc-per-node R tools/banded_nccl_check.py : the row-banded giant world over R GPUs, in both exchange modes
("nccl": torch.distributed collectives; "p2p": CUDA-IPC peer memory + device-side flag barriers, no collective per step).
(1) parity at N = 128 R against the full-torus C oracle (rank 0 gathers the bands); (2) timing at N = 16384."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.distributed as dist
from therldaisyworld_b200.banded import BandedDaisyWorld

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))

MODES = os.environ.get("BANDED_MODES", "nccl,p2p").split(",")
PARTS = os.environ.get("BANDED_PARTS", "parity,timing,fullsize").split(",")


def gather_rows(x):
    t = torch.from_numpy(np.ascontiguousarray(x)).cuda()
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return np.concatenate([o.cpu().numpy() for o in out], axis=1)

# ---- parity
from band_helpers import full_oracle, make_state
N, n, steps = 64 * world * 2, 150, 40
failed = False
for mode, policy in [(m, p) for m in MODES for p in ("none", "replay", "greedy") if "parity" in PARTS]:
    light, dark, ai, st = make_state(N, n, seed=9, clustered=True)
    ai[n // 2:, 0] = (ai[n // 2:, 0] + N // world) % N
    w = BandedDaisyWorld(N, n, rank=rank, world_size=world, device=local, mode=mode)
    w.load_state(light, dark, ai, st)
    ref = full_oracle(w, light, dark, ai, st) if rank == 0 else None       # before the run: takes the clock from w
    acts = np.random.RandomState(2).randint(9, size=(steps, n)) if policy == "replay" else None
    # rank 0 has just spent host time in the CPU oracle: without this barrier the other ranks would already spin in the
    # device-side flag barriers of step 1, whose bounded spin (seconds) turns a slow peer into an error (by design)
    dist.barrier()
    w.run(steps, policy, actions=acts, chunk=16)
    covers = gather_rows(w.local_covers())
    grid = gather_rows(w.local_grid())
    if rank == 0:
        _, done_at, ada = ref.run(steps, policy, actions=None if acts is None else acts[:, None, :])
        checks = {"light": np.array_equal(covers[0], ref.grid[0, 1]), "dark": np.array_equal(covers[1], ref.grid[0, 2]),
                  "grid": np.array_equal(grid, ref.grid[0]), "xy": np.array_equal(w.agents()[0], ref.agent_indices[0]),
                  "state": np.array_equal(w.agents()[1], ref.agent_states[0, :, 0]), "done_at": w.lifespans()[0] == int(done_at[0]),
                  "agents_done_at": np.array_equal(w.lifespans()[1], ada[0, :, 0])}
        ok = all(checks.values())
        print(f"PARITY N={N} ranks={world} {mode} {policy}: {'OK' if ok else 'MISMATCH'} {checks} timed_out={w.band.peer_timed_out()}", flush=True)
        if not ok:
            failed = True
            bad = np.argwhere(covers[0] != ref.grid[0, 1])
            print("  light mismatches:", len(bad), "rows:", sorted(set(bad[:, 0].tolist()))[:20], "cols:", sorted(set(bad[:, 1].tolist()))[:20], flush=True)
            print("  xy mismatches:", int((w.agents()[0] != ref.agent_indices[0]).any(axis=1).sum()), flush=True)
    del w

# ---- timing
flag = torch.tensor([1.0 if failed else 0.0], device="cuda")
dist.all_reduce(flag)
if float(flag[0]) > 0:
    dist.destroy_process_group()
    sys.exit(1)
N, n, K = 16384, 16384, 64
for mode in (MODES if "timing" in PARTS else []):
    w = BandedDaisyWorld(N, n, rank=rank, world_size=world, device=local, mode=mode)
    w.reset_on_device(seed=1)
    w.run(3, "greedy")
    for policy in ("greedy", "none"):
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(K):
            w.step(policy)
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        w.end_chunk()
        t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            ms = float(t[0])
            print(f"TIMING N={N} n={n} ranks={world} {mode} {policy}: {ms / K * 1e3:.1f} us/step (wall {wall / K * 1e6:.1f}) -> "
                  f"{N * N * K / ms / 1e-3:.3e} cell-updates/s timed_out={w.band.peer_timed_out()}", flush=True)
    del w
    torch.cuda.empty_cache()

# ---- full size: R bands vs ONE band (the whole torus on rank 0's GPU), same device-drawn world, bit equality ----------
def checksum(cov):          # exact: milli-cover integers
    k = np.rint(cov * 1000.0).astype(np.int64)
    w = (np.arange(k.shape[-1], dtype=np.int64) % 977 + 1)
    return np.array([k[0].sum(), k[1].sum(), (k[0] * w).sum(), (k[1] * w).sum()], dtype=np.int64)

N, n, K = 16384, 16384, 16
for mode in (MODES if "fullsize" in PARTS else []):
    w = BandedDaisyWorld(N, n, rank=rank, world_size=world, device=local, mode=mode)
    w.reset_on_device(seed=2)
    w.run(K, "greedy", chunk=8)
    cs = w.cover_checksum()                       # device-side, summed over the bands
    ai, st = w.agents()
    life = w.lifespans()
    del w
    torch.cuda.empty_cache()
    if rank == 0:
        one = BandedDaisyWorld(N, n, device=local)
        one.reset_on_device(seed=2)
        one.run(K, "greedy", chunk=8)
        ref_cs = one.cover_checksum()
        ai1, st1 = one.agents()
        life1 = one.lifespans()
        ok = (cs == ref_cs and np.array_equal(ai, ai1) and np.array_equal(st, st1)
              and life[0] == life1[0] and np.array_equal(life[1], life1[1]))
        print(f"FULLSIZE N={N} n={n} {K} greedy steps, {world} bands ({mode}) vs 1 band: {'IDENTICAL' if ok else 'MISMATCH'} "
              f"checksums {cs} vs {ref_cs}", flush=True)
        del one
        torch.cuda.empty_cache()
    dist.barrier()
dist.destroy_process_group()
