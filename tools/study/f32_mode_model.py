"""CPU model of k_forward_f32 (csrc/dw_f32.cuh): the kernel's fp32 arithmetic restated with NumPy float32 (every operation
rounded to binary32, MUFU.RSQ = exact value perturbed by up to 2^-22.9 relative) over whole lives of the NumPy oracle.
Checks (a) that every cell whose bare fraction the screen ACCEPTS rounds like the oracle's b', (b) the fraction of cells
sent to the fp64 tier, (c) the temperature error against the oracle's unrounded fields.
python tools/study/f32_mode_model.py [worlds] [every]"""
import copy, os, sys
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
u = 2.0 ** -24


def F(a, b, c):
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(f32)


def rsq32(x):
    e = (rng.rand(*x.shape) * 2 - 1) * 2.0 ** -22.9
    return ((1.0 / np.sqrt(x.astype(np.float64))) * (1 + e)).astype(f32)


def nsum(k):
    E = np.roll(k, 1, -1) + np.roll(k, -1, -1) + np.roll(k, 1, -2) + np.roll(k, -1, -2)
    C = (np.roll(np.roll(k, 1, -1), 1, -2) + np.roll(np.roll(k, 1, -1), -1, -2) + np.roll(np.roll(k, -1, -1), 1, -2)
         + np.roll(np.roll(k, -1, -1), -1, -2))
    return E, E + C


cells = flagged = bad = 0
maxT = 0.0
maxratio = 0.0
step = 0
while True:
    l, d = env.grid[:, 1].copy(), env.grid[:, 2].copy()
    if max(l.max(), d.max()) <= 0.005 or step > 600:
        break
    act = agent(obs)
    if step > 0 and step % EVERY == 0:
        # the state the forward of this step starts from: after update_agents
        L_now = env.L
        e2 = copy.deepcopy(env)
        e2.update_agents(act)
        l, d = e2.grid[:, 1].copy(), e2.grid[:, 2].copy()
        fl = e2.fields(l, d)                       # unrounded fields of the forward about to happen
        kl, kd = np.rint(l * 1000), np.rint(d * 1000)
        El, Sl = nsum(kl); Ed, Sd = nsum(kd)
        c = env
        dk = np.asarray(c.daisy_kernel)
        w0, w12, w2 = dk[1, 1], dk[0, 1] - dk[0, 0], dk[0, 0]
        a = 1.0 / 8.0
        g2 = c.g * c.g
        cl, cd = (c.albedo_light - c.albedo_bare) / 1000.0, (c.albedo_dark - c.albedo_bare) / 1000.0
        cL = c.S * L_now / c.sigma
        Al0 = c.albedo_bare * c.p; A0 = Al0 * 8 * a
        x0 = g2 * (cL + (c.q - cL) * A0 + (c.q2 - c.q) * Al0 - c.q2 * c.albedo_light)
        xs_l, xs_d = g2 * ((c.q - cL) * a * cl), g2 * ((c.q - cL) * a * cd)
        xk_l, xk_d = g2 * ((c.q2 - c.q) * cl), g2 * ((c.q2 - c.q) * cd)
        xdd = g2 * (c.q2 * (c.albedo_light - c.albedo_dark))
        t0 = -g2 * (c.q2 * (Al0 - c.albedo_light)); tk_l = -g2 * (c.q2 * cl); tk_d = -g2 * (c.q2 * cd)
        topt = np.sqrt(c.g) * c.temp_optimal
        dtp, dtm, dtg = c.dt * c.p, c.dt / 1000.0, c.dt * c.gamma
        k32 = lambda v: v.astype(f32)
        Rl = F(f32(w2), k32(Sl), F(f32(w12), k32(El), (f32(w0) * k32(kl)).astype(f32)))
        Rd = F(f32(w2), k32(Sd), F(f32(w12), k32(Ed), (f32(w0) * k32(kd)).astype(f32)))
        rb = F(f32(-dtm), (Rl + Rd).astype(f32), f32(dtp))
        Xl = (F(f32(xs_l), k32(Sl), F(f32(xs_d), k32(Sd), F(f32(xk_l), k32(kl), (f32(xk_d) * k32(kd)).astype(f32)))) + f32(x0)).astype(f32)
        Xd = (Xl + f32(xdd)).astype(f32)
        XT = (F(f32(tk_l), k32(kl), F(f32(tk_d), k32(kd), f32(t0))) + Xl).astype(f32)
        Tl, Td, T = rsq32(rsq32(Xl)), rsq32(rsq32(Xd)), rsq32(rsq32(XT))
        dTl, dTd = (f32(topt) - Tl).astype(f32), (f32(topt) - Td).astype(f32)
        bl, bd = F(-dTl, dTl, f32(1)), F(-dTd, dTd, f32(1))
        xl = np.clip(F(Rl, F(rb, bl, f32(-dtg)), k32(kl)), f32(0), f32(1000))
        xd = np.clip(F(Rd, F(rb, bd, f32(-dtg)), k32(kd)), f32(0), f32(1000))
        xb = ((f32(1000 * c.p) - xl).astype(f32) - xd).astype(f32)
        kb = np.rint(xb)
        sg, dt = np.sqrt(c.g), abs(c.dt)
        eD = sg * (400.0 * 4.5e-7 + 2 * u * abs(c.temp_optimal))
        rbmax = dt * max(abs(c.p), abs(c.p - 2))
        k1, k2 = f32(2 * 2 * eD * rbmax), f32(2 * u * (17 * rbmax + 8 * dt))
        c0, c0b = f32(2 * u * (2000 + 6000 * dt * abs(c.gamma))), f32(4 * u * 1000 * max(1, abs(c.p)))
        Wl = F(Rl, F(k1, np.abs(dTl), (k2 * (f32(2) - bl)).astype(f32)), c0)
        Wd = F(Rd, F(k1, np.abs(dTd), (k2 * (f32(2) - bd)).astype(f32)), c0)
        ok = (f32(0.5) - np.abs(xb - kb)) > (Wl + Wd + c0b)
        # oracle's b' (unrounded new covers in exact fp64)
        nl = np.clip(l + c.dt * fl["dl"], 0, 1)
        nd = np.clip(d + c.dt * fl["dd"], 0, 1)
        kb_ref = np.rint(((c.p - nl) - nd) * 1000)
        xb_ref = ((c.p - nl) - nd) * 1000
        cells += ok.size; flagged += int((~ok).sum()); bad += int((ok & (kb != kb_ref)).sum())
        err = np.abs(xb.astype(np.float64) - xb_ref)
        maxratio = max(maxratio, float((err / (Wl + Wd + c0b)).max()))
        Tq = T.astype(np.float64) / sg
        maxT = max(maxT, float(np.max(np.abs(Tq - fl["T"]) / fl["T"])))
    obs, *_ = env.step(act)
    step += 1
print(f"steps {step}  cells {cells}  to the fp64 tier {flagged / max(1, cells):.3%}  accepted-but-wrong {bad}  max err/W {maxratio:.3f}  max rel T error {maxT:.2e}")
