#!/usr/bin/env python3
"""Generate golden trajectories from the UNMODIFIED reference (test infrastructure only).

Runs the live reference ``daisy.daisy_world_rl.RLDaisyWorld`` (imported from
/root/reference, which exists only in the build container) together with the
reference ``daisy.agents.greedy.Greedy`` policy, and records full trajectories
into small ``tests/golden/*.npz`` fixtures.  The fixtures travel to the GPU box;
the reference does not.

Usage:  python oracle/gen_golden.py [--ref /root/reference] [--out tests/golden]

Every fixture stores
  meta                json string: constructor kwargs, attribute overrides, seed,
                      policy, number of steps, checkpoint steps
  init_grid           [B,7,N,N] f64   grid right after reset() (unrounded)
  init_agent_indices  [B,n,2]  i64
  init_agent_states   [B,n,1]  f64
  init_obs            [B,n,7,3,3] f64
  actions             [T,b,m,1] i64   action fed to step t (b<=B, m<=n), -1 row => None
  agent_indices       [T,B,n,2] i64   after step t
  agent_states        [T,B,n,1] f64   after step t
  reward, done        [T,...]         as returned by step t
  L                   [T+1] f64       L[t] is the luminosity used in step t
  chan_sum            [T,B,7] f64     per-world per-channel sums of grid after step t
  ckpt_steps          [C] i64         steps (1-based count of steps done) with full dumps
  ckpt_grid           [C,B,7,N,N] f64 grid after those steps
  ckpt_obs            [C,B,n,7,3,3] f64
  diag_temp..         unrounded side-effect attributes after the last step
  done_at, agents_done_at              notebook lifespan counters (cell 2 of
                      notebooks/greedy_longevity_abatement.ipynb) over the T steps
"""
import argparse
import json
import os
import sys
import warnings

import numpy as np


def run_case(name, seed, ctor, attrs, policy, steps, ckpts, out_dir, action_shape=None,
             to_death=False, max_steps=2000):
    from daisy.daisy_world_rl import RLDaisyWorld
    from daisy.agents.greedy import Greedy

    np.random.seed(seed)
    env = RLDaisyWorld(**ctor)
    for k, v in attrs.items():
        if k == "use_microclimate":
            env.set_use_microclimate(v)
        else:
            setattr(env, k, v)
    obs = env.reset()
    rng_state_after_reset = np.random.get_state()

    agent = None
    if policy["kind"] in ("greedy", "antigreedy", "random", "half_random"):
        agent = Greedy()
        agent.greedy = policy["kind"] != "antigreedy"
        agent.epsilon = {"greedy": 0.0, "antigreedy": 0.0, "random": 1.0, "half_random": 0.5}[policy["kind"]]

    mlp_params = None
    if policy["kind"] == "mlp":               # the reference's MLP policy (daisy/agents/mlp.py) with its stored trained weights
        from daisy.agents.mlp import MLP
        import glob
        agent = MLP()
        path = glob.glob(os.path.join(REF_ROOT, "results", "cmaes_exp_002", "*best_agent_gen127.json"))[0]
        mlp_params = np.array(json.load(open(path))["parameters"], dtype=np.float64)
        if policy.get("perturb"):             # a second, different network: the stored one plus seeded noise
            mlp_params = mlp_params * policy.get("scale", 1.0) + np.random.RandomState(policy["perturb"]).randn(mlp_params.size) * policy.get("std", 0.5)
        agent.set_parameters(mlp_params)

    B, n, N = env.batch_size, env.n_agents, env.dim
    rec = dict(
        # global MT19937 state right after reset(): collision fixtures replay the stream from here (actions drawn with
        # randint and the collision noise of step() interleave in it)
        rng_key=rng_state_after_reset[1].copy(), rng_pos=np.int64(rng_state_after_reset[2]),
        init_grid=env.grid.copy(),
        init_agent_indices=env.agent_indices.copy(),
        init_agent_states=env.agent_states.copy(),
        init_obs=obs.copy(),
    )
    T_actions, T_idx, T_st, T_rew, T_done, T_L, T_sum = [], [], [], [], [], [env.L], []
    ck_steps, ck_grid, ck_obs = [], [], []
    done_at = np.zeros((B,), dtype=np.int64)
    agents_done_at = np.zeros((B, n, 1), dtype=np.int64)
    t = 0
    while True:
        if policy["kind"] == "none":
            action = None
        elif policy["kind"] == "fixed":       # cycles through a list of actions of a given shape
            a = policy["cycle"][t % len(policy["cycle"])]
            action = np.full(action_shape, a, dtype=np.int64)
        elif policy["kind"] == "randint":     # fresh uniform 0..8 per agent from the global stream
            action = np.random.randint(9, size=(B, n, 1))
        else:
            action = agent(obs)
        if action is None:
            T_actions.append(np.full((1, 1, 1), -1, dtype=np.int64))
        else:
            T_actions.append(np.asarray(action, dtype=np.int64).copy())
        obs, reward, done, info = env.step(action)
        t += 1
        T_idx.append(env.agent_indices.copy())
        T_st.append(env.agent_states.copy())
        T_rew.append(np.asarray(reward).copy())
        T_done.append(np.asarray(done).copy())
        T_L.append(env.L)
        T_sum.append(env.grid.sum(axis=(-2, -1)))
        grid_done = env.grid[:, 1:3, :, :].max(axis=(1, 2, 3)) <= 0.005
        done_at += (1 - 1 * grid_done)
        if n:
            agents_done_at += (1 - 1 * done)
        if t in ckpts or (to_death and grid_done.all()) or (not to_death and t == steps):
            if t not in ck_steps:
                ck_steps.append(t)
                ck_grid.append(env.grid.copy())
                ck_obs.append(obs.copy())
        if to_death:
            if grid_done.mean() == 1.0 or t >= max_steps:
                break
        elif t == steps:
            break

    same_shape = all(a.shape == T_actions[0].shape for a in T_actions)
    assert same_shape
    rec.update(
        actions=np.stack(T_actions),
        agent_indices=np.stack(T_idx), agent_states=np.stack(T_st),
        reward=np.stack(T_rew), done=np.stack(T_done), L=np.array(T_L),
        chan_sum=np.stack(T_sum),
        ckpt_steps=np.array(ck_steps, dtype=np.int64),
        ckpt_grid=np.stack(ck_grid), ckpt_obs=np.stack(ck_obs),
        diag_temp=env.temp.copy(), diag_temp_light=env.temp_light.copy(),
        diag_temp_dark=env.temp_dark.copy(), diag_temp_effective=env.temp_effective.copy(),
        diag_dead_temp=env.dead_temp.copy(), diag_beta=env.beta.copy(),
        diag_beta_l=env.beta_l.copy(), diag_beta_d=env.beta_d.copy(), diag_growth=env.growth.copy(),
        done_at=done_at, agents_done_at=agents_done_at,
    )
    if mlp_params is not None:
        rec["mlp_params"] = mlp_params
    meta = dict(name=name, seed=seed, ctor=ctor, attrs=attrs, policy=policy, steps=t,
                to_death=to_death, B=B, n=n, N=N, numpy=np.__version__,
                final_step_count=int(env.step_count), dL=float(env.dL))
    rec["meta"] = np.array(json.dumps(meta))
    path = os.path.join(out_dir, name + ".npz")
    np.savez_compressed(path, **rec)
    print(f"{name}: B={B} n={n} N={N} steps={t} done_at[:4]={done_at[:4]} "
          f"-> {path} ({os.path.getsize(path)/1024:.0f} KiB)")


REF_ROOT = "/root/reference"

CASES = [
    # next row N1: the reference's MLP policy (63-16-32-9 ReLU, argmax) with the stored trained weights / a perturbed copy
    dict(name="mlp_n16_b4_200", seed=17, ctor=dict(grid_dimension=16), attrs=dict(batch_size=4),
         policy=dict(kind="mlp"), steps=200, ckpts=[1, 2, 100]),
    dict(name="mlp_n64_b2_n6_60", seed=19, ctor=dict(grid_dimension=64, n_agents=6), attrs=dict(batch_size=2),
         policy=dict(kind="mlp", perturb=2, std=2.0), steps=60, ckpts=[1, 30]),
    dict(name="mlp_n16_b4_mixed_150", seed=17, ctor=dict(grid_dimension=16), attrs=dict(batch_size=4),
         policy=dict(kind="mlp", perturb=2, std=2.0), steps=150, ckpts=[1, 75]),
    # BASELINE config 1: default grid, no agents, single world, run to biosphere death (seed 42)
    dict(name="cfg1_n16_b1_noagents_todeath", seed=42, ctor=dict(n_agents=0), attrs=dict(batch_size=1),
         policy=dict(kind="none"), steps=0, ckpts=[1, 2, 100, 300, 440], to_death=True),
    # greedy light/dark, N=16, full life (small version of config 2)
    dict(name="greedy_n16_b4_todeath", seed=13, ctor=dict(grid_dimension=16), attrs=dict(batch_size=4),
         policy=dict(kind="greedy"), steps=0, ckpts=[1, 2, 3, 50, 200, 400], to_death=True),
    # README experiment shape (N=8), anti-greedy
    dict(name="antigreedy_n8_b16_todeath", seed=13, ctor=dict(grid_dimension=8), attrs=dict(batch_size=16),
         policy=dict(kind="antigreedy"), steps=0, ckpts=[1, 10, 100, 300], to_death=True),
    # stochastic policies -> replayed actions
    dict(name="random_n8_b8_todeath", seed=7, ctor=dict(grid_dimension=8), attrs=dict(batch_size=8),
         policy=dict(kind="random"), steps=0, ckpts=[1, 64, 256], to_death=True),
    dict(name="halfrandom_n16_b4_300", seed=5, ctor=dict(grid_dimension=16), attrs=dict(batch_size=4),
         policy=dict(kind="half_random"), steps=300, ckpts=[1, 150]),
    # config-2 shape at reduced batch: 64x64 greedy light/dark
    dict(name="greedy_n64_b2_120", seed=13, ctor=dict(grid_dimension=64), attrs=dict(batch_size=2),
         policy=dict(kind="greedy"), steps=120, ckpts=[1, 60]),
    # neutral albedo (config 3 condition), 64x64 anti-greedy, short
    dict(name="neutral_antigreedy_n64_b1_40", seed=21, ctor=dict(grid_dimension=64),
         attrs=dict(batch_size=1, albedo_light=0.5, albedo_dark=0.5),
         policy=dict(kind="antigreedy"), steps=40, ckpts=[1]),
    # odd sizes the FFT path supports (5, 7, 17), many agents, no microclimate, other constants
    dict(name="greedy_n5_b3_n2_todeath", seed=3, ctor=dict(grid_dimension=5, n_agents=2), attrs=dict(batch_size=3),
         policy=dict(kind="greedy"), steps=0, ckpts=[1, 2, 100], to_death=True),
    dict(name="randint_n7_b3_n16_200", seed=9, ctor=dict(grid_dimension=7, n_agents=16), attrs=dict(batch_size=3),
         policy=dict(kind="randint"), steps=200, ckpts=[1, 2, 3, 100]),
    dict(name="greedy_n17_b2_params_200", seed=11, ctor=dict(grid_dimension=17, n_agents=3, ramp_period=128),
         attrs=dict(batch_size=2, use_microclimate=False, dt=0.5, agent_gamma=0.03, gamma=0.2,
                    min_L=0.8, max_L=1.7, albedo_bare=0.45, albedo_light=0.8, albedo_dark=0.2,
                    initial_al=0.3, initial_ad=0.1, light_proportion=0.5, dark_proportion=0.25),
         policy=dict(kind="greedy"), steps=200, ckpts=[1, 2, 100]),
    # action-shape quirks (A.4-7): (1,1,1) action while B=32,n=4, cycling 0..8 (reference's own test)
    dict(name="fixed_smallaction_n16_b32_27", seed=1, ctor=dict(), attrs=dict(),
         policy=dict(kind="fixed", cycle=list(range(9))), action_shape=(1, 1, 1), steps=27, ckpts=[1, 9]),
    dict(name="fixed_partial_n16_b5_n4_36", seed=2, ctor=dict(), attrs=dict(batch_size=5),
         policy=dict(kind="fixed", cycle=[8, 5, 4, 7, 6, 0, 3]), action_shape=(3, 2, 1), steps=36, ckpts=[1, 7]),
    # step(None) with agents: drift + starvation (A.4-2)
    dict(name="none_n16_b2_n4_40", seed=4, ctor=dict(), attrs=dict(batch_size=2),
         policy=dict(kind="none"), steps=40, ckpts=[1, 18, 19]),
    # N4 row: collision_mode == 1 (reference :220-242). Crowded small grids so several agents share a cell every few steps.
    dict(name="collide_randint_n5_b3_n12_120", seed=23, ctor=dict(grid_dimension=5, n_agents=12, collision_mode=1),
         attrs=dict(batch_size=3), policy=dict(kind="randint"), steps=120, ckpts=[1, 2, 60]),
    dict(name="collide_greedy_n8_b4_n10_150", seed=29, ctor=dict(grid_dimension=8, n_agents=10, collision_mode=1),
         attrs=dict(batch_size=4, food_chain_penalty=0.8), policy=dict(kind="greedy"), steps=150, ckpts=[1, 75]),
    dict(name="collide_randint_n7_b2_n40_60", seed=31, ctor=dict(grid_dimension=7, n_agents=40, collision_mode=1),
         attrs=dict(batch_size=2, agent_gamma=0.01), policy=dict(kind="randint"), steps=60, ckpts=[1, 30]),
    # A15: non-default observation masks (daisy/nn/functional.py:51-103). Greedy reads only the four edge cells, the MLP all 63
    # inputs, so with "moore" the corner cells decide actions
    dict(name="mlp_moore_n16_b3_100", seed=37, ctor=dict(grid_dimension=16, neighborhood_mode="moore"), attrs=dict(batch_size=3),
         policy=dict(kind="mlp", perturb=23, std=2.0), steps=100, ckpts=[1, 2, 50]),
    dict(name="greedy_moore_n64_b2_n6_60", seed=39, ctor=dict(grid_dimension=64, n_agents=6, neighborhood_mode="moore"),
         attrs=dict(batch_size=2), policy=dict(kind="greedy"), steps=60, ckpts=[1, 30]),
    dict(name="mlp_circular_n16_b2_60", seed=37, ctor=dict(grid_dimension=16, neighborhood_mode="circular"), attrs=dict(batch_size=2),
         policy=dict(kind="mlp", perturb=23, std=2.0, scale=0.3), steps=60, ckpts=[1, 30]),
    # ramp_up_down branch (N4 row, cheap to pin): short ramp so dL flips sign
    dict(name="rampupdown_n8_b2_100", seed=6, ctor=dict(grid_dimension=8, ramp_period=32),
         attrs=dict(batch_size=2, ramp_up_down=True, ddL=0.01),
         policy=dict(kind="greedy"), steps=100, ckpts=[1, 32, 33, 64]),
]


def record_masks(out_dir):
    """make_neighborhood of the live reference (daisy/nn/functional.py:51-103) for every mode and radius 1..3."""
    from daisy.nn.functional import make_neighborhood
    rec = {}
    for kr in (1, 2, 3):
        for mode in ("moore", "von_neumann", "circular", "nonsense"):
            rec[f"{mode}_{kr}"] = make_neighborhood(radius=kr, mode=mode)
    rec["default"] = make_neighborhood()
    path = os.path.join(out_dir, "ref_masks.npz")
    np.savez_compressed(path, **rec)
    print(f"masks -> {path}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(os.path.dirname(__file__), "..", "tests", "golden"))
    ap.add_argument("--only", default=None)
    args = ap.parse_args()
    global REF_ROOT
    REF_ROOT = args.ref
    sys.path.insert(0, args.ref)
    warnings.filterwarnings("ignore", category=DeprecationWarning)
    os.makedirs(args.out, exist_ok=True)
    if not args.only or args.only == "masks":
        record_masks(args.out)
    for case in CASES:
        if args.only and args.only not in case["name"]:
            continue
        c = dict(case)
        run_case(c.pop("name"), c.pop("seed"), c.pop("ctor"), c.pop("attrs"), c.pop("policy"),
                 c.pop("steps"), c.pop("ckpts"), args.out, **c)


if __name__ == "__main__":
    main()
