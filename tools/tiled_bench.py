"""Time the giant-grid path in isolation on one GPU: N x N torus, n agents, K steps after the literal first step."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from therldaisyworld_b200.banded import BandedDaisyWorld

def bench(N, n, policy, K=64, reps=3):
    w = BandedDaisyWorld(N, n)
    w.reset_on_device(seed=1)
    w.run(2, policy)
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        w.band.run_local(K, policy)
        e1.record()
        torch.cuda.synchronize()
        w._pending += K
        w.end_chunk()
        best = min(best, e0.elapsed_time(e1))
    print(f"N={N} n={n} {policy}: {best / K * 1e3:.1f} us/step -> {N * N * K / best / 1e-3:.3e} cell-updates/s, "
          f"HBM {8 * N * N * K / best / 1e-3 / 1e9:.0f} GB/s algorithmic, slow cells {w.band.slow_count()}", flush=True)

if __name__ == "__main__":
    if len(sys.argv) > 1:                       # profiling target: one size, few steps
        N = int(sys.argv[1])
        bench(N, N, "greedy", K=8, reps=1)
        sys.exit(0)
    for N, n in ((1024, 1024), (4096, 4096), (16384, 16384), (16384, 0)):
        bench(N, n, "greedy")
