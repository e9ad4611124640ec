#!/usr/bin/env python3
"""Benchmark of the RLDaisyWorld simulation step (BASELINE.json metric: cell-updates/s & env-steps/s, batched 64x64).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (config.workload): BASELINE configs[1] -- a 1000-world ensemble (per GPU), 64x64, light(0.75)/dark(0.25)
daisies, 4 greedy agents per world, seed 13.  One bench "step" = one pass of the hot path over that batch: the
whole ensemble advanced by T=384 env steps from the post-reset state (every world is still alive at step 384, so
no work is skipped): 1 literal materialising step (the reset state is off the 0.001 lattice) + 383 fused steps.
`value` = cell-updates/s = worlds * 64*64 * T * K / (device time, max over ranks), state resident in HBM.
`e2e` = the same through the C-ABI with HOST buffers: per step, upload of the two cover planes + agents from pinned
host memory, the T-step run, and the download of the lifespan counters, all inside the timed region.
N > 1: one process per GPU (torchrun), worlds sharded (1000 per rank, weak scaling), no data-path collective;
one NCCL all-reduce of the 8-double lifespan-statistics vector per bench step, inside the timed region.

The same JSON line carries, under "giant_grid", a second measured workload (not the headline): BASELINE configs[4], one
16384x16384 toroidal world with 16384 greedy agents, row-banded over the N GPUs (strong scaling; halo rows + two small
all-reduces per step over NCCL), timed with CUDA events, max over ranks; under "sustained" seconds-long whole-life
ensembles (configs[2] shape at N = 1, configs[3] at N > 1) with the clocks sampled; under "dropin_step" single env.step()
calls through the Python drop-in.  --no-extras skips the three.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly ONE JSON line: keep a private handle to it and point fd 1 at stderr, so that anything a library
# prints (NCCL writes its version banner to stdout) cannot end up in front of the line.
_JSON_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)

import numpy as np  # noqa: E402

WORLDS = 1000
N = 64
N_AGENTS = 4
T_STEPS = 384
SEED = 13
FLOP_PER_CELL_UPDATE = 98          # algorithmic flops of one cell-update
WORKLOAD = (f"BASELINE configs[1]: {WORLDS}-world ensemble per GPU, {N}x{N}, light/dark daisies, {N_AGENTS} greedy "
            f"agents, seed {SEED}; one step = {T_STEPS} env steps from the reset state")


# ----------------------------------------------------------------------------------------------- reference arm
REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def have_reference():
    """The UNMODIFIED reference installed by baseline/install_ref.py (git-ignored, travels with the snapshot)."""
    return os.path.exists(os.path.join(REF_DIR, "daisy", "nn", "functional.py"))


def make_cpu_env(worlds, seed):
    """(env, policy, kind): the reference's own RLDaisyWorld + Greedy from baseline/_ref when present ("reference"), else
    the NumPy port with the same FFT convolutions and Python agent loops ("port")."""
    import warnings
    warnings.filterwarnings("ignore")
    np.random.seed(seed)
    if have_reference():
        if REF_DIR not in sys.path:
            sys.path.insert(0, REF_DIR)
        from daisy.daisy_world_rl import RLDaisyWorld as RefWorld
        from daisy.agents.greedy import Greedy
        env = RefWorld(grid_dimension=N, n_agents=N_AGENTS)
        agent, kind = Greedy(), "reference"
    else:
        from oracle.daisy_numpy import OracleDaisyWorld, OracleGreedy
        env = OracleDaisyWorld(conv="fft", grid_dimension=N, n_agents=N_AGENTS)
        agent, kind = OracleGreedy(), "port"
    env.batch_size = worlds
    return env, agent, kind


def _ref_worker(args):
    """One host process: its share of the ensemble stepped by the reference's CPU implementation."""
    rank, worlds, steps_per_sample, n_samples, warm = args
    os.environ["OMP_NUM_THREADS"] = "1"
    env, agent, kind = make_cpu_env(worlds, SEED + rank)
    obs = env.reset()
    out = []
    for s in range(warm + n_samples):
        t0 = time.perf_counter()
        for _ in range(steps_per_sample):
            obs, _, _, _ = env.step(agent(obs))
        out.append(time.perf_counter() - t0)
    return out[warm:], kind


def run_reference(args):
    """--impl reference: the reference's own NumPy CPU path (the unmodified reference from baseline/_ref when it travelled,
    else the oracle port), all host cores, on a bounded sample of the same workload."""
    import multiprocessing as mp
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 32, WORLDS))
    per = [WORLDS // procs + (1 if i < WORLDS % procs else 0) for i in range(procs)]
    steps_per_sample = 2                       # env steps per bench step: bounded sample (1000 worlds x 2 steps)
    K, W = max(1, args.steps), max(0, args.warmup)
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(procs) as pool:
        res = pool.map(_ref_worker, [(i, per[i], steps_per_sample, K, W) for i in range(procs)])
    times, kind = [r[0] for r in res], res[0][1]
    per_step = np.max(np.array(times), axis=0)          # slowest process bounds each step
    total = float(per_step.sum())
    cells = WORLDS * N * N * steps_per_sample * K
    value = cells / total
    sample = (f"{WORLDS} worlds x {steps_per_sample} env steps per bench step, split over {procs} processes; "
              + ("unmodified reference RLDaisyWorld + Greedy from baseline/_ref" if kind == "reference"
                 else "NumPy port (oracle/daisy_numpy.py, FFT convolutions): baseline/_ref absent"))
    line = {
        "impl": "reference", "metric": "cell_updates_per_s", "value": value, "unit": "cell-updates/s", "n_gpus": args.gpus,
        "steps": K, "warmup": W, "ms_per_step": 1e3 * total / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "env_steps_per_s": WORLDS * steps_per_sample * K / total,
        "cpu_baseline": {"value": value, "unit": "cell-updates/s", "cores": procs, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line), file=_JSON_OUT, flush=True)


# ----------------------------------------------------------------------------------------------- helpers
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, power, reasons = [], [], [], set()
        try:
            for ln in open(self.path):
                f = [x.strip() for x in ln.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            # "under load": samples in the upper half of the observed power range
            thr = (min(power) + max(power)) / 2 if power else 0
            load = [s for s, p in zip(sm, power) if p >= thr] or sm
            out.update(sm_mhz=float(np.median(load)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm),
                       power_w_max=float(max(power)))
        return out


def initial_state(rank):
    """Post-reset state of this rank's ensemble, produced exactly like RLDaisyWorld.reset() draws it
    (daisy_world_rl.py:285-302, 173-179) -- synthetic data, seed SEED + rank."""
    rng = np.random.RandomState(SEED + rank)
    rng.randint(N, size=(32, N_AGENTS, 2))                # the constructor's extra draw
    u_dark = rng.rand(WORLDS, 2, N, N)
    u_light = rng.rand(WORLDS, 2, N, N)
    dark = 1.0 * (u_dark[:, 0] < 0.33) * 0.2 * u_dark[:, 1]
    light = 1.0 * (u_light[:, 0] < 0.33) * 0.2 * u_light[:, 1]
    agents = rng.randint(N, size=(WORLDS, N_AGENTS, 2)).astype(np.int64)
    states = np.ones((WORLDS, N_AGENTS))
    return np.ascontiguousarray(light), np.ascontiguousarray(dark), agents, states


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return None


def load_traffic():
    """DRAM bytes per launch of the fused kernel from the committed ncu capture (profiles/), or None."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "fused_traffic.json"))).get("dram_bytes_per_launch")
    except Exception:
        return None


def cpu_baseline_sample():
    """The reference's CPU path (unmodified reference from baseline/_ref when present, else the NumPy port), 1 core,
    bounded sample."""
    env, agent, kind = make_cpu_env(WORLDS, SEED)
    obs = env.reset()
    obs, _, _, _ = env.step(agent(obs))                   # warm-up
    steps = 5
    t0 = time.perf_counter()
    for _ in range(steps):
        obs, _, _, _ = env.step(agent(obs))
    dt = time.perf_counter() - t0
    what = "unmodified reference (baseline/_ref)" if kind == "reference" else "NumPy port with FFT convolutions"
    return {"value": WORLDS * N * N * steps / dt, "unit": "cell-updates/s", "cores": 1, "kind": kind,
            "sample": f"{WORLDS} worlds x {N}x{N} x {steps} env steps after 1 warm-up, {what}",
            "env_steps_per_s": WORLDS * steps / dt, "seconds": dt}


# ----------------------------------------------------------------------------------------------- sustained ensembles
def sustained_bench(rank, world, local):
    """Seconds-long runs on the fused path, device-timed, clocks sampled (the headline's timed region is ~0.1 s).
    N = 1: BASELINE configs[2] shape -- 100 000 worlds of 64x64 with 4 agents, WHOLE LIVES under the notebook's stopping rule
    (32-step segments run past the stopping step, surplus trimmed out of the agents' counters: ensemble.simulate_lifespan),
    four of its conditions back to back (>= 2 s of device time).
    N > 1: BASELINE configs[3] -- 125 000 worlds per GPU (10^6 at N = 8), random agents, whole lives, worlds sharded with no
    data-path collective; per 64-step segment one MIN all-reduce of the all-done flags, at the end ONE NCCL all-reduce of the
    8-double lifespan statistics."""
    import torch
    import torch.distributed as dist
    from therldaisyworld_b200 import RLDaisyWorld
    from therldaisyworld_b200.ensemble import DeviceShard, shard_range, simulate_lifespan
    total = 100_000 if world == 1 else 125_000 * world
    lo, hi = shard_range(total, world, rank)
    np.random.seed(0)
    env = RLDaisyWorld(grid_dimension=N, n_agents=N_AGENTS, device=local)
    env.batch_size = hi - lo
    conditions = ([("light/dark", "greedy"), ("light/dark", "random"), ("neutral", "greedy"), ("neutral", "antigreedy")]
                  if world == 1 else [("light/dark", "random")])
    sampler = ClockSampler(local) if rank == 0 else None
    out, dev_ms, cells = [], 0.0, 0
    for albedo, policy in conditions:
        env.albedo_light, env.albedo_dark = (0.75, 0.25) if albedo == "light/dark" else (0.5, 0.5)
        env.reset_on_device(seed=SEED, world_offset=lo)
        env.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = simulate_lifespan(DeviceShard(env, world_offset=lo), policy=policy, seed=7, device="cuda")
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
        c = r["worlds"] * N * N * r["steps"]
        dev_ms += ms
        cells += c
        out.append({"albedo": albedo, "policy": policy, "worlds": r["worlds"], "steps_to_last_death": r["steps"], "device_ms": ms,
                    "cell_updates_per_s": c / (ms * 1e-3), "env_steps_per_s": r["worlds"] * r["steps"] / (ms * 1e-3),
                    "biosphere_lifespan": [r["biosphere_lifespan_mean"], r["biosphere_lifespan_sem"]],
                    "agent_lifespan": [r.get("agent_lifespan_mean"), r.get("agent_lifespan_sem")]})
    clocks = sampler.stop() if sampler else None
    del env
    torch.cuda.empty_cache()
    return {"workload": (f"BASELINE configs[2] shape: {total} worlds x {N}x{N}, {N_AGENTS} agents, whole lives (stopping rule: "
                         "32-step segments, surplus steps trimmed out of the lifespan counters), 4 conditions") if world == 1 else
                        (f"BASELINE configs[3]: {total} worlds ({total // world} per GPU) x {N}x{N}, random agents, whole lives, "
                         "one NCCL all-reduce of the 8-double statistics + one MIN all-reduce per 64-step segment"),
            "metric": "cell_updates_per_s", "value": cells / (dev_ms * 1e-3), "unit": "cell-updates/s", "n_gpus": world,
            "device_seconds": dev_ms * 1e-3, "scaling": "weak", "conditions": out, "clocks": clocks,
            "timing": "CUDA events around each condition (incl. the first literal step, the segment read-backs and the trim), max over ranks"}


# ----------------------------------------------------------------------------------------------- fp32 mode
def materialise_bench(local, peaks, worlds=4000, reps=10):
    """Device time of materialising env.grid of a lattice-resident ensemble: fp64 (k_forward_x2 + stamp) vs the fp32 mode
    (k_forward_f32 + stamp, fp32 arithmetic from the packed lattice). Algorithmic HBM bytes per cell: fp64 4 (lattice) + 7 * 8;
    fp32 4 + 4 (the two lattices) + 7 * 4."""
    from therldaisyworld_b200 import RLDaisyWorld
    np.random.seed(SEED)
    env = RLDaisyWorld(grid_dimension=N, n_agents=N_AGENTS, device=local)
    env.batch_size = worlds
    env.reset_on_device(seed=SEED)
    env.run(40, policy="greedy")
    lib, h = env._lib, env._h
    out = {"worlds": worlds, "grid": N, "reps": reps}
    hbm = peaks.get("hbm_gbs") if peaks else 6650.0
    for name, flag, bytes_per_cell in (("fp64", 0, 4 + 7 * 8), ("fp32", 1, 8 + 7 * 4)):
        ms = C.c_double()
        _lib_check(lib, h, lib.dw_debug_time_materialise(h, flag, reps, C.byref(ms)), "dw_debug_time_materialise")
        gbs = worlds * N * N * bytes_per_cell / (ms.value * 1e-3) / 1e9
        out[name] = {"ms": ms.value, "cell_updates_per_s": worlds * N * N / (ms.value * 1e-3), "algorithmic_bytes_per_cell": bytes_per_cell,
                     "hbm_gbs": gbs, "hbm_frac": gbs / hbm}
    st = env.f32_stats()
    out["fp32"]["cells_to_fp64_tier"] = st["fp64_tier"] / max(1, st["cells"])
    out["fp32"]["cells_to_literal"] = st["literal"] / max(1, st["cells"])
    out["hbm_peak_gbs"] = hbm
    out["what"] = ("env.grid of a lattice-resident ensemble materialised on the device, CUDA events; fp32: covers/bare fraction exact "
                   "(as binary32), temperatures in fp32 arithmetic (1e-5 tolerance of the north_star's fp32 mode)")
    del env
    return out


def mlp_policy_bench(local, worlds=WORLDS, steps=T_STEPS - 1, reps=3):
    """Row N1 at the bench's ensemble shape: the reference's MLP policy (63-16-32-9, random weights) chosen on the device every
    step. 64x64 worlds: windows + network inside the persistent fused kernel; wall time of env.run incl. the final sync."""
    from therldaisyworld_b200 import RLDaisyWorld
    np.random.seed(SEED)
    env = RLDaisyWorld(grid_dimension=N, n_agents=N_AGENTS, device=local)
    env.batch_size = worlds
    env.reset()
    env.set_mlp(np.random.RandomState(0).randn(1808))
    env.run(1, policy="mlp")
    lib, h = env._lib, env._h
    _lib_check(lib, h, lib.dw_checkpoint_save(h), "dw_checkpoint_save")
    best = None
    for _ in range(reps):
        _lib_check(lib, h, lib.dw_checkpoint_restore(h), "dw_checkpoint_restore")
        env._pull_clock()
        env.synchronize()
        t0 = time.perf_counter()
        env.run(steps, policy="mlp")
        env.synchronize()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    del env
    return {"worlds": worlds, "grid": N, "n_agents": N_AGENTS, "env_steps": steps, "ms": best * 1e3,
            "cell_updates_per_s": worlds * N * N * steps / best, "env_steps_per_s": worlds * steps / best,
            "what": "env.run(policy='mlp'): MLP.get_action (daisy/agents/mlp.py:97-116) evaluated inside the fused kernel, wall clock"}


def _lib_check(lib, h, rc, what):
    from therldaisyworld_b200 import _lib
    _lib.check(lib, h, rc, what)


# ----------------------------------------------------------------------------------------------- giant grid (configs[4])
GIANT_N = 16384
GIANT_STEPS = 64


def giant_grid_bench(rank, world, local, fp64_peak_tflops):
    """One 16384^2 world, 16384 greedy agents, row-banded over `world` GPUs. Every rank calls this (collectives inside)."""
    import torch
    import torch.distributed as dist
    from therldaisyworld_b200.banded import BandedDaisyWorld
    mode = "local"
    if world > 1:
        # peer-memory mode (CUDA IPC + device-side barriers, no collective per step); NCCL mode if the mapping fails anywhere
        mode = "p2p"
        try:
            w = BandedDaisyWorld(GIANT_N, GIANT_N, rank=rank, world_size=world, device=local, mode="p2p")
            ok = 1.0
        except Exception:
            w, ok = None, 0.0
        flag = torch.tensor([ok], dtype=torch.float64, device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if float(flag[0]) < 1.0:
            mode = "nccl"
            del w
            w = BandedDaisyWorld(GIANT_N, GIANT_N, rank=rank, world_size=world, device=local, mode="nccl")
    else:
        w = BandedDaisyWorld(GIANT_N, GIANT_N, rank=rank, world_size=world, device=local)
    w.reset_on_device(seed=SEED)
    w.run(3, "greedy")                                   # literal first step + 2 lattice steps (warm-up)
    best = None
    for _ in range(2):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if world == 1:
            w.band.run_local(GIANT_STEPS, "greedy")
            w._pending += GIANT_STEPS
        else:
            for _ in range(GIANT_STEPS):
                w.step("greedy")
        e1.record()
        torch.cuda.synchronize()
        w.end_chunk()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        best = float(t[0]) if best is None else min(best, float(t[0]))
    cells = GIANT_N * GIANT_N * GIANT_STEPS
    value = cells / (best * 1e-3)
    done_at, ada = w.lifespans()
    # state checksums after the 3 + 2 x GIANT_STEPS steps: exact integers, independent of the banding -> equal at N = 1/2/4/8
    cover_cs = w.cover_checksum()
    ai, st = w.agents()
    agent_cs = [int((ai[:, 0].astype(np.int64) * 16411 + ai[:, 1]).sum()), float(np.sum(st)), int(np.sum(ada))]
    # compute floor of a step: this band's stencil kernel alone, back to back (max over ranks); the rest of a step is the
    # agent kernels, the halo push, the barriers and launch gaps
    ts = torch.tensor([w.band.time_stencil(20)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ts, op=dist.ReduceOp.MAX)
    us_stencil = float(ts[0])
    out = {
        "workload": f"BASELINE configs[4]: single {GIANT_N}x{GIANT_N} toroidal world, {GIANT_N} greedy agents, row-banded over "
                    f"{world} GPU(s) ({GIANT_N // world} rows each), {GIANT_STEPS} steps after 3 warm-up steps",
        "metric": "cell_updates_per_s", "value": value, "unit": "cell-updates/s", "scaling": "strong", "n_gpus": world,
        "us_per_step": best * 1e3 / GIANT_STEPS, "us_stencil": us_stencil, "us_exchange": best * 1e3 / GIANT_STEPS - us_stencil,
        "env_steps_per_s": GIANT_STEPS / (best * 1e-3),
        "hbm_gbs_algorithmic_per_gpu": 8.0 * value / world / 1e9,
        "fp64_roofline_frac": value * FLOP_PER_CELL_UPDATE / 1e12 / (fp64_peak_tflops * world) if fp64_peak_tflops else None,
        "exchange_mode": mode,
        "exchanges_per_step": {"local": "none (toroidal wrap on the device)",
                               "p2p": "owner-publishes P2P stores of decisions/gains into every rank's exchange vector, edge rows pushed "
                                      "into the neighbours' ghost rows, 2 device-side flag barriers; no collective. Two streams: edge "
                                      "tiles, halo push, barriers and the look-ahead decisions of the next step run on a high-priority "
                                      "side stream under the interior tiles",
                               "nccl": "1 all-reduce(SUM, 2n doubles) + 1 ring halo exchange overlapped with the interior tiles"}[mode],
        "peer_barrier_timed_out": bool(w.band.peer_timed_out()) if mode == "p2p" else None,
        "literal_recomputations": w.band.slow_count(), "biosphere_alive_steps": done_at,
        "state_checksum": {"steps": 3 + 2 * GIANT_STEPS, "covers": cover_cs, "agents": agent_cs,
                           "what": "covers: dwt_cover_checksum summed over the bands (sum kl, sum kd, position-weighted sums, mod 2^64); "
                                   "agents: [sum(x * 16411 + y), sum(state), sum(agents_done_at)] -- identical for every number of bands"},
    }
    del w
    torch.cuda.empty_cache()
    return out


# ----------------------------------------------------------------------------------------------- product arm
def run_product(args):
    import torch
    import torch.distributed as dist
    from therldaisyworld_b200 import RLDaisyWorld, _lib
    from therldaisyworld_b200._lib import DwProfile, DwRunResult, DW_POLICY

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    K, W = max(1, args.steps), max(3, args.warmup)

    light, dark, agents, states = initial_state(rank)
    # the drop-in object gives us a configured handle (constants, kernels, clock); state comes from initial_state()
    np.random.seed(SEED + rank)
    env = RLDaisyWorld(grid_dimension=N, n_agents=N_AGENTS, device=local)
    env.batch_size = WORLDS
    env.reset()
    lib, h = env._lib, env._h
    lib.dw_set_world_offset(h, rank * WORLDS)

    def chk(rc, what):
        _lib.check(lib, h, rc, what)

    pd = C.POINTER(C.c_double)
    # pinned host staging for the e2e arm
    pin_light = torch.from_numpy(light).pin_memory()
    pin_dark = torch.from_numpy(dark).pin_memory()
    pin_agents = torch.from_numpy(agents).pin_memory()
    pin_states = torch.from_numpy(states).pin_memory()
    out_done_at = torch.zeros(WORLDS, dtype=torch.int64).pin_memory()
    out_agents = torch.zeros(WORLDS, N_AGENTS, dtype=torch.int64).pin_memory()
    stats = torch.zeros(8, dtype=torch.float64, device="cuda")

    def upload():
        chk(lib.dw_upload_covers(h, C.cast(pin_light.data_ptr(), pd), C.cast(pin_dark.data_ptr(), pd)), "dw_upload_covers")
        chk(lib.dw_upload_state(h, None, C.cast(pin_agents.data_ptr(), C.POINTER(C.c_int64)), C.cast(pin_states.data_ptr(), pd)),
            "dw_upload_state")
        env.L, env.step_count = env.min_L, 0
        env.dL = (env.max_L - env.min_L) / env.ramp_period
        clk = env._clock()
        chk(lib.dw_set_clock(h, C.byref(clk)), "dw_set_clock")
        chk(lib.dw_reset_lifespans(h), "dw_reset_lifespans")

    res = DwRunResult()

    def run_T():
        chk(lib.dw_run(h, T_STEPS, DW_POLICY["greedy"], None, C.c_uint64(0), 0, C.byref(res)), "dw_run")
        chk(lib.dw_lifespan_stats_device(h, C.c_void_p(stats.data_ptr())), "dw_lifespan_stats_device")
        if world > 1:
            dist.all_reduce(stats)

    def download():
        chk(lib.dw_get_lifespans(h, C.cast(out_done_at.data_ptr(), C.POINTER(C.c_int64)),
                                 C.cast(out_agents.data_ptr(), C.POINTER(C.c_int64))), "dw_get_lifespans")

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident arm: state already in HBM (device-side checkpoint), timed with CUDA events on the launch stream
    upload()
    chk(lib.dw_checkpoint_save(h), "dw_checkpoint_save")
    for _ in range(W):
        chk(lib.dw_checkpoint_restore(h), "dw_checkpoint_restore")
        run_T()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    chk(lib.dw_set_profiling(h, 1), "dw_set_profiling")
    t_res = 0.0
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    wall0 = time.perf_counter()
    for k in range(K):
        chk(lib.dw_checkpoint_restore(h), "dw_checkpoint_restore")
        flush.fill_(k & 0xff)                                    # L2 flush between timed iterations
        ev[k][0].record()
        run_T()
        ev[k][1].record()
    barrier()
    wall_res = time.perf_counter() - wall0
    t_res = sum(a.elapsed_time(b) for a, b in ev) * 1e-3
    prof = DwProfile()
    chk(lib.dw_get_profile(h, C.byref(prof)), "dw_get_profile")
    chk(lib.dw_set_profiling(h, 0), "dw_set_profiling")
    download()
    mean_life = float(out_done_at.double().mean())               # all alive at T => T_STEPS exactly

    # ---- e2e arm: host buffers in, host results out, every step.  Two handles on two streams, used alternately: the
    # upload of step k+1 (65 MB over PCIe from pinned memory) is issued before the (blocking) run of step k, so the copy
    # engine works under the compute of the previous step -- what a throughput-minded caller of the C-ABI would do.  Every
    # step's host->device copy, run and device->host read are inside the timed region.
    np.random.seed(SEED + rank)
    env_b = RLDaisyWorld(grid_dimension=N, n_agents=N_AGENTS, device=local)
    env_b.batch_size = WORLDS
    env_b.reset()
    lib.dw_set_world_offset(env_b._h, rank * WORLDS)
    slots = []
    for e in (env, env_b):
        st = torch.cuda.Stream()
        chk(lib.dw_set_stream(e._h, C.c_void_p(st.cuda_stream)), "dw_set_stream")
        slots.append({"env": e, "h": e._h, "stream": st, "stats": torch.zeros(8, dtype=torch.float64, device="cuda"),
                      "done_at": torch.zeros(WORLDS, dtype=torch.int64).pin_memory(),
                      "agents": torch.zeros(WORLDS, N_AGENTS, dtype=torch.int64).pin_memory()})
    torch.cuda.synchronize()

    def e2e_upload(s):
        e, hh = s["env"], s["h"]
        chk(lib.dw_upload_covers(hh, C.cast(pin_light.data_ptr(), pd), C.cast(pin_dark.data_ptr(), pd)), "dw_upload_covers")
        chk(lib.dw_upload_state(hh, None, C.cast(pin_agents.data_ptr(), C.POINTER(C.c_int64)), C.cast(pin_states.data_ptr(), pd)),
            "dw_upload_state")
        e.L, e.step_count = e.min_L, 0
        e.dL = (e.max_L - e.min_L) / e.ramp_period
        clk = e._clock()
        chk(lib.dw_set_clock(hh, C.byref(clk)), "dw_set_clock")
        chk(lib.dw_reset_lifespans(hh), "dw_reset_lifespans")

    def e2e_run_download(s):
        hh = s["h"]
        chk(lib.dw_run(hh, T_STEPS, DW_POLICY["greedy"], None, C.c_uint64(0), 0, C.byref(res)), "dw_run")
        chk(lib.dw_lifespan_stats_device(hh, C.c_void_p(s["stats"].data_ptr())), "dw_lifespan_stats_device")
        if world > 1:
            with torch.cuda.stream(s["stream"]):
                dist.all_reduce(s["stats"])
        chk(lib.dw_get_lifespans(hh, C.cast(s["done_at"].data_ptr(), C.POINTER(C.c_int64)),
                                 C.cast(s["agents"].data_ptr(), C.POINTER(C.c_int64))), "dw_get_lifespans")

    def e2e_loop(steps):
        e2e_upload(slots[0])
        for k in range(steps):
            if k + 1 < steps:
                e2e_upload(slots[(k + 1) % 2])
            e2e_run_download(slots[k % 2])
        torch.cuda.synchronize()

    e2e_loop(3)
    barrier()
    flush.fill_(1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e2e_loop(K)
    t_e2e = time.perf_counter() - t0
    barrier()
    clocks = sampler.stop() if sampler else None
    out_done_at, out_agents = slots[(K - 1) % 2]["done_at"], slots[(K - 1) % 2]["agents"]
    mean_life_e2e = float(out_done_at.double().mean())
    for s_ in slots:
        chk(lib.dw_set_stream(s_["h"], None), "dw_set_stream")

    # ---- max over ranks
    t = torch.tensor([t_res, t_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t_res, t_e2e = float(t[0]), float(t[1])
    cells_per_step = world * WORLDS * N * N * T_STEPS
    value = cells_per_step * K / t_res
    e2e_value = cells_per_step * K / t_e2e
    h2d = world * (light.nbytes + dark.nbytes + agents.nbytes + states.nbytes)       # whole job, all ranks
    d2h = world * (out_done_at.numel() * 8 + out_agents.numel() * 8)

    # measured FP64 FMA peak of this device (the fused kernel's roofline denominator)
    tf, ms = C.c_double(), C.c_double()
    chk(lib.dw_debug_fp64_peak(h, 20000, 5, C.byref(tf), C.byref(ms)), "dw_debug_fp64_peak")
    giant = None
    if not args.no_extras:
        try:
            giant = giant_grid_bench(rank, world, local, tf.value)
        except Exception as e:                                    # the headline line must survive
            giant = {"error": repr(e)}

    sustained = None
    if not args.no_extras:
        try:
            sustained = sustained_bench(rank, world, local)
        except Exception as e:
            sustained = {"error": repr(e)}
    dropin = None
    if rank == 0 and not args.no_extras:
        # the path every reference caller uses: ONE env.step() per step through the Python drop-in (lattice-resident state,
        # K = 1 launch of the fused kernel, one packed device->host copy). Secondary record, not the headline.
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            from dropin_bench import measure
            dropin = measure(WORLDS, N, N_AGENTS, iters=200, warm=20, device=local)
            dropin["what"] = ("single env.step() calls through the drop-in class at the bench's ensemble shape: host Greedy policy "
                              "(obs download + action upload every step), device policy with / without the observation download")
        except Exception as e:
            dropin = {"error": repr(e)}
    materialise = None
    if rank == 0 and not args.no_extras:
        # fp32 mode: the 7-channel grid of a lattice-resident state materialised with fp32 arithmetic (k_forward_f32) against the
        # fp64 materialisation (k_forward_x2), both HBM-bound; 4000 worlds so that the output (0.46 / 0.92 GB) exceeds the L2
        try:
            materialise = materialise_bench(local, load_peaks())
        except Exception as e:
            materialise = {"error": repr(e)}
    mlp_rec = None
    if rank == 0 and not args.no_extras:
        try:
            mlp_rec = mlp_policy_bench(local)
        except Exception as e:
            mlp_rec = {"error": repr(e)}
    if rank == 0:
        peaks = load_peaks()
        fused_s = prof.fused_ms * 1e-3
        achieved = prof.fused_cell_updates * FLOP_PER_CELL_UPDATE / fused_s / 1e12 if fused_s > 0 else None
        roofline = {
            "bound": "fp64_fma", "kernel": "k_fused (SMEM-resident lattice kernel)",
            "achieved": achieved, "peak": tf.value, "unit": "TFLOP/s",
            "frac": (achieved / tf.value) if achieved else None,
            # second denominator: the nominal FP64 FMA rate, 148 SMs x 64 DFMA/clk x 2 flop x 1.965 GHz (max SM clock)
            "peak_nominal": 148 * 64 * 2 * 1.965e9 / 1e12,
            "frac_nominal": (achieved / (148 * 64 * 2 * 1.965e9 / 1e12)) if achieved else None,
            "peak_source": "measured live: dw_debug_fp64_peak (dependent DFMA chains, full occupancy); MEASURED_PEAKS.json "
                           "has no FP64 entry",
            "flop_per_cell_update": FLOP_PER_CELL_UPDATE,
            "cell_updates_per_launch": prof.fused_cell_updates / max(1, prof.fused_launches),
            "avg_launch_ms": prof.fused_ms / max(1, prof.fused_launches),
            "kernel_share_of_step": fused_s / t_res if t_res > 0 else None,
            "kernel_cell_updates_per_s": prof.fused_cell_updates / fused_s if fused_s > 0 else None,
            "traffic": load_traffic(),
            "hbm": {"algorithmic_bytes_per_launch": 8 * WORLDS * N * N * ((T_STEPS - 1 + 15) // 16),
                    "peak_gbs": peaks.get("hbm_gbs") if peaks else 6650.0,
                    "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback",
                    "note": "packed lattice, 4 B in + 4 B out per cell per 16-step chunk (L2-resident between chunks): "
                            "the kernel is not HBM bound; `traffic` is the DRAM bytes per launch from profiles/fused_traffic.json"},
        }
        cpu = cpu_baseline_sample() if world == 1 else None
        line = {
            "metric": "cell_updates_per_s", "value": value, "unit": "cell-updates/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": 1e3 * t_res / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "worlds_per_gpu": WORLDS, "grid": N, "n_agents": N_AGENTS, "policy": "greedy",
                       "env_steps_per_bench_step": T_STEPS,
                       "l2": "resident arm: flushed between timed iterations (256 MB write); e2e arm: inputs arrive from the host each step",
                       "e2e_pipeline": "two handles on two streams: upload of step k+1 overlaps the run of step k",
                       "parallelism": f"worlds sharded over {world} GPU(s), no data-path collective"},
            "env_steps_per_s": world * WORLDS * T_STEPS * K / t_res,
            "e2e": {"value": e2e_value, "unit": "cell-updates/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * t_e2e / K, "env_steps_per_s": world * WORLDS * T_STEPS * K / t_e2e},
            "gpu_launches": int(prof.kernel_launches),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "clocks": clocks,
            "giant_grid": giant,
            "sustained": sustained,
            "dropin_step": dropin,
            "fp32_mode": materialise,
            "mlp_policy": mlp_rec,
            "check": {"mean_done_at_after_T": mean_life, "mean_done_at_after_T_e2e": mean_life_e2e, "expected": float(T_STEPS),
                      "ensemble_stats": stats.tolist(),
                      "wall_s_resident": wall_res},
        }
        print(json.dumps(line), file=_JSON_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary giant-grid measurement")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_product(args)


if __name__ == "__main__":
    main()
