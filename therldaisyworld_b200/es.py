"""ES fitness rollouts on the device (next-row N2): ``SimpleGaussianES.get_fitness`` (``daisy/evo/sges.py:144-181``) for a
whole population in one batch.

The reference evaluates its members one after the other (or one per MPI rank, ``sges.py:299-349``): each call resets the
32-world environment and lets the member's MLP drive the first half of every world's agents against an adversary MLP on
the second half, summing ``reward[:, :half].mean()`` per step until all agents are done or ``max_steps``.  Here the
members' environments are laid side by side as ONE batch of ``P x worlds_per_member`` worlds: observations, both networks
(per-world weights), the step and the fitness bookkeeping all run on the device; the host only draws the reset states.

The resets are drawn from the global NumPy stream in exactly the order the sequential reference loop would draw them
(member 0's dark, light, agents, then member 1's, ...), so with the same seed the fitness values reproduce the reference's.
"""
import ctypes as C

import numpy as np

from ._lib import DwRunResult  # noqa: F401  (keeps the binding module loaded)
from .env import RLDaisyWorld, _ptr


def evaluate_population(members, adversary_idx=0, max_steps=768, worlds_per_member=32, env=None, member_range=None,
                        device_reset_seed=None, **env_kwargs):
    """members: [P, 1808] MLP parameter vectors (MLP.get_parameters()). Returns (fitness[P], total_steps[P, W, n, 1],
    member_steps[P], env). Pass `env` (an RLDaisyWorld of this package) to reuse a handle / non-default constants.
    member_range=(lo, hi): evaluate only those members (the reset draws of all P are still consumed, so every rank of a
    sharded evaluation sees the stream the sequential reference loop would).
    device_reset_seed: draw the members' reset states on the device (counter RNG keyed by the world index: same
    distribution as the reference's reset, NOT its NumPy stream) instead of on the host -- the host draws in the
    reference's RNG order cost ~12 ms per generation of 64 members; for training runs that do not need the reference's
    exact stream."""
    all_members = np.ascontiguousarray(np.asarray(members, dtype=np.float64))
    lo, hi = (0, all_members.shape[0]) if member_range is None else member_range
    W = int(worlds_per_member)
    if device_reset_seed is not None:
        if member_range is not None:
            raise ValueError("device_reset_seed is for single-rank evaluations (use member_range with the host draws)")
        return _evaluate(all_members, adversary_idx, max_steps, W, env, env_kwargs, device_seed=int(device_reset_seed))
    if member_range is not None and hi <= lo:
        raise ValueError("empty member range (more ranks than members?)")
    if member_range is not None:
        # local population = my members + the adversary (appended when it is not one of mine)
        idx = list(range(lo, hi))
        if adversary_idx not in idx:
            idx.append(adversary_idx)
        local_adv = idx.index(adversary_idx)
        fitness, total_steps, member_steps, env = _evaluate(all_members[idx], local_adv, max_steps, W, env, env_kwargs,
                                                            skip=(lo, all_members.shape[0] - hi), n_eval=hi - lo, draw_index=idx)
        return fitness[:hi - lo], total_steps[:hi - lo], member_steps[:hi - lo], env
    return _evaluate(all_members, adversary_idx, max_steps, W, env, env_kwargs)


def evaluate_population_sharded(members, adversary_idx=0, max_steps=768, worlds_per_member=32, group=None, device=None, **env_kwargs):
    """The mantle/arm fan-out of the reference (sges.py:299-349) without MPI: every rank of a torch.distributed group
    evaluates a contiguous slice of the population on its GPU, then the fitness vector is all-gathered. Returns fitness[P]."""
    import torch
    import torch.distributed as dist
    from .ensemble import shard_range
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    P = np.asarray(members).shape[0]
    lo, hi = shard_range(P, world, rank)
    if device is not None:
        env_kwargs["device"] = device
    fit, _, _, _ = evaluate_population(members, adversary_idx, max_steps, worlds_per_member, member_range=(lo, hi), **env_kwargs)
    full = torch.zeros(P, dtype=torch.float64, device="cuda" if dist.get_backend(group) == "nccl" else "cpu")
    full[lo:hi] = torch.from_numpy(fit)
    dist.all_reduce(full, group=group)          # disjoint slices: the sum is a gather
    return full.cpu().numpy()


def _evaluate(members, adversary_idx, max_steps, W, env, env_kwargs, skip=(0, 0), n_eval=None, draw_index=None, device_seed=None):
    P = members.shape[0]
    if env is None:
        state = np.random.get_state()
        env = RLDaisyWorld(**env_kwargs)               # the constructor's own draws must not disturb the caller's stream
        np.random.set_state(state)
    N, n = int(env.dim), int(env.n_agents)
    B = P * W
    if device_seed is not None:
        env.batch_size = B
        env.reset_on_device(seed=device_seed)
        return _rollout(env, members, adversary_idx, max_steps, P, W, n, B)
    # the P resets of the sequential reference loop, in its RNG order (daisy_world_rl.py:285-302, 173-179)
    light = np.empty((B, N, N))
    dark = np.empty((B, N, N))
    agents = np.empty((B, n, 2), dtype=np.int64)
    def draw():
        dp = np.random.rand(W, 2, N, N)
        lp = np.random.rand(W, 2, N, N)
        return (1.0 * (dp[:, 0] < env.dark_proportion) * env.initial_ad * dp[:, 1],
                1.0 * (lp[:, 0] < env.light_proportion) * env.initial_al * lp[:, 1], np.random.randint(N, size=(W, n, 2)))

    n_eval = P if n_eval is None else n_eval
    for _ in range(skip[0]):                       # members evaluated by lower ranks: consume their draws
        draw()
    for m in range(P):
        if m < n_eval:
            dark[m * W:(m + 1) * W], light[m * W:(m + 1) * W], agents[m * W:(m + 1) * W] = draw()
        else:                                      # the appended adversary block is never scored: any valid state will do
            dark[m * W:(m + 1) * W], light[m * W:(m + 1) * W], agents[m * W:(m + 1) * W] = dark[:W], light[:W], agents[:W]
    for _ in range(skip[1]):
        draw()
    states = np.ones((B, n))
    env.batch_size = B
    env.L = env.min_L
    env.dL = (env.max_L - env.min_L) / env.ramp_period
    env.step_count = 0
    env._ensure_handle((B, N, n))
    for mm in env._m.values():
        mm.invalidate()
    env._push()
    lib, h = env._lib, env._h
    env._check(lib.dw_upload_covers(h, _ptr(light, C.c_double), _ptr(dark, C.c_double)), "dw_upload_covers")
    env._check(lib.dw_upload_state(h, None, _ptr(agents, C.c_int64), _ptr(states, C.c_double)), "dw_upload_state")
    env._check(lib.dw_init_temperatures(h), "dw_init_temperatures")
    return _rollout(env, members, adversary_idx, max_steps, P, W, n, B)


def _rollout(env, members, adversary_idx, max_steps, P, W, n, B):
    lib, h = env._lib, env._h
    env._check(lib.dw_set_mlp_population(h, _ptr(members, C.c_double), P, int(adversary_idx)), "dw_set_mlp_population")
    steps = C.c_int64(0)
    env._check(lib.dw_run_population(h, int(max_steps), C.byref(steps)), "dw_run_population")
    env._state_changed()
    env._pull_clock()
    fitness = np.zeros(P)
    member_steps = np.zeros(P, dtype=np.int64)
    total_steps = np.zeros((B, n), dtype=np.int64)
    env._check(lib.dw_get_population_results(h, _ptr(fitness, C.c_double), _ptr(member_steps, C.c_int64), _ptr(total_steps, C.c_int64)),
               "dw_get_population_results")
    return fitness, total_steps.reshape(P, W, n, 1), member_steps, env
