"""CPU study for the fp32 screening tier of the lattice fast path: run the NumPy oracle over whole lives,
evaluate every live cell with (a) the fp64 fast-path formulas and (b) their fp32 restatement (every operation rounded to
binary32, MUFU.RSQ modelled as the correctly rounded value perturbed by up to 2^-22.9 relative), and report
  * the largest |x32 - x64| (milli-cover) against candidate per-cell bounds,
  * the fraction of cells a tie filter of that width would send on to the fp64 tier.
python tools/study/fp32_tier_study.py [worlds] [every]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from oracle.daisy_numpy import OracleDaisyWorld, OracleGreedy

f32 = np.float32
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
EVERY = int(sys.argv[2]) if len(sys.argv) > 2 else 7
N = 64
np.random.seed(13)
env = OracleDaisyWorld(grid_dimension=N, n_agents=4)
env.batch_size = B
obs = env.reset()
agent = OracleGreedy()
rng = np.random.RandomState(5)

def fma32(a, b, c):       # one rounding, like FFMA
    return (a.astype(np.float64) * np.float64(b) + np.float64(c)).astype(f32) if np.isscalar(b) or np.isscalar(c) else \
        (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)

def F(a, b, c):
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(f32)

def rsq32(x):             # MUFU.RSQ model: exact value, relative perturbation up to 2^-22.9, rounded to fp32
    e = (rng.rand(*x.shape) * 2 - 1) * 2.0 ** -22.9
    return ((1.0 / np.sqrt(x.astype(np.float64))) * (1 + e)).astype(f32)

w = np.array(env.daisy_kernel if hasattr(env, "daisy_kernel") else None)
stats = {"n": 0, "max_err": 0.0, "max_ratio": {}, "fall": {}}
errs, bounds = [], {}
step = 0
tot_cells = 0
fall = {}
maxratio = {}
while True:
    l, d = env.grid[:, 1], env.grid[:, 2]
    if max(l.max(), d.max()) <= 0.005 or step > 600:
        break
    if step > 0 and step % EVERY == 0:
        kl = np.rint(l * 1000); kd = np.rint(d * 1000)
        def nsum(k):
            E = np.roll(k, 1, -1) + np.roll(k, -1, -1) + np.roll(k, 1, -2) + np.roll(k, -1, -2)
            C = (np.roll(np.roll(k, 1, -1), 1, -2) + np.roll(np.roll(k, 1, -1), -1, -2) + np.roll(np.roll(k, -1, -1), 1, -2)
                 + np.roll(np.roll(k, -1, -1), -1, -2))
            return E, E + C
        El, Sl = nsum(kl); Ed, Sd = nsum(kd)
        c = env
        dk = np.asarray(c.daisy_kernel if hasattr(c, "daisy_kernel") else c.kernel_daisy)
        w0, w12, w2 = dk[1, 1], dk[0, 1] - dk[0, 0], dk[0, 0]
        a = 1.0 / 8.0
        g2 = c.g * c.g
        cl, cd = (c.albedo_light - c.albedo_bare) / 1000.0, (c.albedo_dark - c.albedo_bare) / 1000.0
        cL = c.S * c.L / c.sigma
        Al0 = c.albedo_bare * c.p; A0 = Al0 * 8 * a
        x0 = g2 * (cL + (c.q - cL) * A0 + (c.q2 - c.q) * Al0 - c.q2 * c.albedo_light)
        xs_l, xs_d = g2 * ((c.q - cL) * a * cl), g2 * ((c.q - cL) * a * cd)
        xk_l, xk_d = g2 * ((c.q2 - c.q) * cl), g2 * ((c.q2 - c.q) * cd)
        xdd = g2 * (c.q2 * (c.albedo_light - c.albedo_dark))
        topt = np.sqrt(c.g) * c.temp_optimal
        dtp, dtm, dtg = c.dt * c.p, c.dt / 1000.0, c.dt * c.gamma
        # fp64 fast path
        Rl = w2 * Sl + (w12 * El + w0 * kl); Rd = w2 * Sd + (w12 * Ed + w0 * kd)
        rb = -dtm * (Rl + Rd) + dtp
        Xl = xs_l * Sl + (xs_d * Sd + (xk_l * kl + (xk_d * kd + x0))); Xd = Xl + xdd
        dTl = topt - Xl ** 0.25; dTd = topt - Xd ** 0.25
        xl64 = Rl * (rb * (1 - dTl * dTl) - dtg) + kl
        xd64 = Rd * (rb * (1 - dTd * dTd) - dtg) + kd
        # fp32 path
        k32 = lambda v: v.astype(f32)
        c32 = lambda v: f32(v)
        Rl3 = F(c32(w2), k32(Sl), F(c32(w12), k32(El), (c32(w0) * k32(kl)).astype(f32)))
        Rd3 = F(c32(w2), k32(Sd), F(c32(w12), k32(Ed), (c32(w0) * k32(kd)).astype(f32)))
        rb3 = F(c32(-dtm), (Rl3 + Rd3).astype(f32), c32(dtp))
        Xl3 = F(c32(xs_l), k32(Sl), F(c32(xs_d), k32(Sd), F(c32(xk_l), k32(kl), F(c32(xk_d), k32(kd), c32(x0)))))
        Xd3 = (Xl3 + c32(xdd)).astype(f32)
        Tl3 = rsq32(rsq32(Xl3)); Td3 = rsq32(rsq32(Xd3))
        dTl3 = (c32(topt) - Tl3).astype(f32); dTd3 = (c32(topt) - Td3).astype(f32)
        bl3 = F(-dTl3, dTl3, f32(1)); bd3 = F(-dTd3, dTd3, f32(1))
        xl32 = F(Rl3, F(rb3, bl3, c32(-dtg)), k32(kl)); xd32 = F(Rd3, F(rb3, bd3, c32(-dtg)), k32(kd))
        live = (kl + kd + Sl + Sd) > 0
        for nm, x64, x32, R3, dT3 in (("l", xl64, xl32, Rl3, dTl3), ("d", xd64, xd32, Rd3, dTd3)):
            err = np.abs(x32.astype(np.float64) - x64)[live]
            Rr = np.abs(R3.astype(np.float64) * rb3)[live]; aT = np.abs(dT3.astype(np.float64))[live]; R_ = R3.astype(np.float64)[live]
            x = x64[live]
            frac = np.abs((x + 0.5) - np.rint(x + 0.5))          # distance from a rounding tie
            cands = {
                "const_0.04": np.full_like(err, 0.04),
                "c*|dT|+c0": 5e-3 * aT + 8e-4,
                "c*R*rb*|dT|+c0": 2.0e-5 * Rr * aT + 8e-4,
                "c*R*rb*|dT|+c1*R+c0": 1.7e-5 * Rr * aT + 4e-7 * R_ + 5e-4,
            }
            for cn, bnd in cands.items():
                key = cn
                maxratio[key] = max(maxratio.get(key, 0.0), float((err / bnd).max()))
                fall.setdefault(key, [0, 0])
                fall[key][0] += int((frac < bnd).sum()); fall[key][1] += err.size
            stats["max_err"] = max(stats["max_err"], float(err.max()))
        tot_cells += int(live.sum())
    obs, *_ = env.step(agent(obs))
    step += 1
print("steps", step, "live cells sampled", tot_cells, "max |x32-x64| milli", stats["max_err"])
for k in maxratio:
    print(f"{k:28s} max err/bound {maxratio[k]:.3f}   fall-through per species {fall[k][0] / fall[k][1]:.4%}")
