"""Randomised stress test of the lattice fast path + tie filter: random physics constants, the fused kernels against the
literal materialising kernels (DW_DISABLE_FUSED=1 DW_LITERAL_ONLY=1) on the same worlds. Any mismatch = a fast-path result further from the
literal value than the tie filter assumes. Usage: python tools/fuzz_fast_path.py [n_configs] [seed] [sizes, e.g. 8,16,32,48,96,128]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from therldaisyworld_b200 import RLDaisyWorld


def draw(rng, sizes=(64, 64, 16, 33, 96)):
    a = dict(albedo_bare=rng.uniform(0.3, 0.7), S=rng.uniform(700, 1300), g=10 ** rng.uniform(-3, -2), gamma=rng.uniform(0.05, 0.45),
             dt=rng.choice([0.25, 0.5, 1.0]), temp_optimal=rng.uniform(280, 310), agent_gamma=rng.uniform(0.01, 0.1),
             min_L=rng.uniform(0.5, 0.9), initial_al=rng.uniform(0.05, 0.6), initial_ad=rng.uniform(0.05, 0.6),
             light_proportion=rng.uniform(0.1, 0.9), dark_proportion=rng.uniform(0.1, 0.9))
    a["albedo_light"] = a["albedo_bare"] + rng.uniform(0.0, 0.3)
    a["albedo_dark"] = a["albedo_bare"] - rng.uniform(0.0, 0.3)
    a["max_L"] = a["min_L"] + rng.uniform(0.3, 1.2)
    return a, int(rng.choice(list(sizes))), int(rng.choice([0, 1, 4, 9])), int(rng.choice([64, 128, 512])), bool(rng.rand() < 0.2)


def run_config(attrs, N, n, ramp, no_micro, seed, B=24, steps=400, policy="greedy"):
    out = []
    for disable in (False, True):
        if disable:                                  # the ground truth: materialising kernels, literal arithmetic only
            os.environ["DW_DISABLE_FUSED"] = "1"
            os.environ["DW_LITERAL_ONLY"] = "1"
        else:
            os.environ.pop("DW_DISABLE_FUSED", None)
            os.environ.pop("DW_LITERAL_ONLY", None)
        try:
            np.random.seed(seed)
            env = RLDaisyWorld(grid_dimension=N, n_agents=n, ramp_period=ramp)
            env.batch_size = B
            for k, v in attrs.items():
                setattr(env, k, v)
            env.q = 0.2 * env.S / env.sigma
            env.set_use_microclimate(not no_micro)
            env.reset()
            env.reset_lifespans()
            env.run(steps, policy=policy)
            cnt = C.c_uint64()
            env._lib.dw_debug_slow_count(env._h, C.byref(cnt), 0)
            out.append((env.grid[:, 1:3].copy(), env.agent_states.copy(), env.lifespans()[0].copy(), cnt.value))
        finally:
            os.environ.pop("DW_DISABLE_FUSED", None)
            os.environ.pop("DW_LITERAL_ONLY", None)
    (g0, s0, l0, slow), (g1, s1, l1, _) = out
    ok = np.array_equal(g0, g1) and np.array_equal(s0, s1) and np.array_equal(l0, l1)
    return ok, slow, int((g0 != g1).sum()), float(l0.mean())


if __name__ == "__main__":
    n_cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    rng = np.random.RandomState(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    sizes = tuple(int(v) for v in sys.argv[3].split(",")) if len(sys.argv) > 3 else (64, 64, 16, 33, 96)
    bad = 0
    for c in range(n_cfg):
        attrs, N, n, ramp, no_micro = draw(rng, sizes)
        ok, slow, diff, life = run_config(attrs, N, n, ramp, no_micro, seed=c)
        cells = 24 * N * N * 400
        print(f"cfg {c}: N={N} n={n} ramp={ramp} micro={not no_micro} mean life {life:.0f}: {'OK' if ok else 'MISMATCH'} diff_cells={diff} "
              f"literal recomputations {slow} ({slow / cells:.1e})", flush=True)
        if not ok:
            bad += 1
            print("   attrs:", {k: (float(v) if not isinstance(v, str) else v) for k, v in attrs.items()}, flush=True)
    print("FUZZ", "PASSED" if bad == 0 else f"FAILED ({bad} configs)")
    sys.exit(1 if bad else 0)
