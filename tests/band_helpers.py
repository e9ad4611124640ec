"""Helpers shared by the giant-grid (row-banded) tests: synthetic states, the full-torus oracle, an in-process
communicator that lets several bands of ONE process step in lock-step (threads), and an oracle-backed band."""
import threading

import numpy as np


def make_state(N, n, seed, clustered=False):
    """Off-lattice reset state with the reference's distribution (daisy_world_rl.py:285-302); optionally agents packed
    into a 5x5 patch straddling the (0,0) corner so that grazing collisions and toroidal wraps happen every step."""
    rng = np.random.RandomState(seed)
    u = rng.rand(2, 2, N, N)
    dark = 1.0 * (u[0, 0] < 0.33) * 0.2 * u[0, 1]
    light = 1.0 * (u[1, 0] < 0.33) * 0.2 * u[1, 1]
    if clustered:
        ai = (rng.randint(5, size=(n, 2)) - 2) % N
    else:
        ai = rng.randint(N, size=(n, 2))
    return light, dark, ai.astype(np.int64), np.ones(n)


def full_oracle(world, light, dark, ai, st):
    """C oracle of the whole torus as one batch element, with the world's constants and clock."""
    from oracle.daisy_c import COracleWorld
    N = light.shape[0]
    grid = np.zeros((1, 7, N, N))
    grid[0, 1], grid[0, 2] = light, dark
    return COracleWorld(world, grid=grid, agent_indices=ai[None].copy(), agent_states=st[None, :, None].copy())


class ThreadComm:
    """banded.DistComm stand-in for several bands living in one process: every band runs in its own thread and the
    exchanges are barrier-synchronised tensor copies (works for CUDA tensors on one device and for CPU tensors)."""

    class Shared:
        def __init__(self, world_size):
            self.barrier = threading.Barrier(world_size)
            self.slots = [None] * world_size

    def __init__(self, rank, world_size, shared):
        self.rank, self.world_size, self.sh = rank, world_size, shared

    def _reduce(self, t, op):
        import torch
        sh = self.sh
        sh.slots[self.rank] = t
        sh.barrier.wait()
        if t.is_cuda:                # bands of one process may run on different streams: order them device-wide (test only)
            torch.cuda.synchronize()
        stacked = torch.stack(list(sh.slots))
        out = stacked.sum(0) if op == "sum" else stacked.max(0).values
        sh.barrier.wait()            # everyone has read every slot
        t.copy_(out)
        sh.barrier.wait()

    def all_reduce_sum(self, t):
        self._reduce(t, "sum")

    def all_reduce_max(self, t):
        self._reduce(t, "max")

    def exchange_halos(self, send_top, send_bottom, recv_top, recv_bottom):
        sh = self.sh
        up, down = (self.rank - 1) % self.world_size, (self.rank + 1) % self.world_size
        sh.slots[self.rank] = (send_top, send_bottom)
        sh.barrier.wait()
        if send_top.is_cuda:      # bands of one process may exchange on side streams: order them device-wide (test only)
            import torch
            torch.cuda.synchronize()
        recv_top.copy_(sh.slots[up][1])
        recv_bottom.copy_(sh.slots[down][0])
        sh.barrier.wait()


def run_threads(worlds, fn):
    """fn(world) for every band concurrently; re-raises the first exception."""
    errs = []

    def wrap(w):
        try:
            fn(w)
        except BaseException as e:      # noqa: BLE001
            errs.append(e)
            try:
                w.comm.sh.barrier.abort()
            except Exception:
                pass

    # Garbage from earlier tests (banded worlds sit in reference cycles) must not be collected while bands are stepping:
    # freeing a handle calls cudaFree, which waits for the device to drain -- but a band spinning in a peer barrier only
    # finishes once THIS thread has launched its next kernels, so the collection would stall until the spin times out.
    import gc
    gc.collect()
    gc.disable()
    try:
        ts = [threading.Thread(target=wrap, args=(w,)) for w in worlds]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
    finally:
        gc.enable()
    if errs:
        raise errs[0]


class OracleBand:
    """CPU stand-in for banded.DeviceBand (test infrastructure): the same phase interface on NumPy arrays, arithmetic
    from the NumPy oracle (oracle/daisy_numpy.py) applied to the band padded with its ghost rows."""

    def __init__(self, params, N, n_agents, row0, rows, n_ranks):
        import torch
        from oracle.daisy_numpy import OracleDaisyWorld
        self._torch = torch
        self.N, self.n, self.row0, self.rows, self.n_ranks = N, n_agents, row0, rows, n_ranks
        self.o = OracleDaisyWorld(grid_dimension=N, n_agents=0)
        self.set_params(params)
        self.exch = torch.zeros(2 * n_agents, dtype=torch.float64)      # [gain | act] like the device band
        self.gain = self.exch[:n_agents]
        self.act = self.exch[n_agents:]
        self.stepmax = torch.zeros(4096 * 2, dtype=torch.int32)
        self.j = 0
        self.done_at = 0

    def set_params(self, p):
        for k in ("p", "g", "S", "sigma", "gamma", "q", "q2", "temp_optimal", "dt", "agent_gamma", "albedo_bare", "albedo_light",
                  "albedo_dark"):
            setattr(self.o, k, getattr(p, k))

    def set_clock(self, clk):
        import copy
        self.clk = copy.copy(clk)

    def get_clock(self):
        return self.clk

    def upload(self, light_rows, dark_rows, agent_indices, agent_states):
        self.l = np.array(light_rows, dtype=np.float64)       # [rows+2, N]
        self.d = np.array(dark_rows, dtype=np.float64)
        self.xy = np.array(agent_indices, dtype=np.int64).reshape(self.n, 2)
        self.st = np.array(agent_states, dtype=np.float64).reshape(self.n)
        self.ada = np.zeros(self.n, dtype=np.int64)
        self.gz = np.zeros(self.n, dtype=bool)
        self.j = 0

    def _images(self, x):
        rel = (x - self.row0) % self.N
        out = []
        if rel < self.rows:
            out.append(rel + 1)
        if rel == self.N - 1:
            out.append(0)
        if rel == (0 if self.rows == self.N else self.rows):
            out.append(self.rows + 1)
        return out

    def _owned(self, x):
        rel = (x - self.row0) % self.N
        return rel + 1 if rel < self.rows else -1

    def decide(self, policy, actions_step=None, seed=0):
        self._absorb_halo()
        N = self.N
        food = self.l + self.d
        for i in range(self.n):
            x, y = self.xy[i]
            if policy == "replay":
                self.act[i] = float(int(np.asarray(actions_step).reshape(-1)[i]) + 1)
            elif policy == "none":
                self.act[i] = 1.0
            else:
                lr = self._owned(x)
                if lr < 0:
                    self.act[i] = 0.0
                    continue
                f = [food[lr, (y - 1) % N], food[lr - 1, y], food[lr + 1, y], food[lr, (y + 1) % N]]
                k = int(np.argmax(f)) if policy == "greedy" else int(np.argmin(f))
                self.act[i] = float(4 + k + 1)

    def move_graze(self):
        N = self.N
        self.st = self.st - self.o.agent_gamma
        self.gz[:] = False
        gains = np.zeros(self.n)
        for i in range(self.n):          # index order: the first grazer of a cell eats, later ones find it empty
            if not self.st[i] > 0.0:
                continue
            a = int(self.act[i].item()) - 1
            if a != 8:
                axis, step = ((1, -1), (0, -1), (0, 1), (1, 1))[a % 4]
                self.xy[i, axis] = (self.xy[i, axis] + step) % N
            if a > 4:
                self.gz[i] = True
                x, y = self.xy[i]
                lr = self._owned(x)
                if lr >= 0:
                    gains[i] = self.l[lr, y] + self.d[lr, y]
                for r in self._images(x):
                    self.l[r, y] = 0.0
                    self.d[r, y] = 0.0
        self.gain[:] = self._torch.from_numpy(gains)

    def finish_agents(self):
        g = self.gain.numpy()
        s = np.where(self.gz, self.st + g, self.st)
        self.st = np.clip(s, 0.0, 1.0)
        self.ada += (self.st >= 0.1)

    def stencil(self, part=0):
        from oracle.daisy_numpy import round3
        self.o.L = self.clk.L
        f = self.o.fields(self.l[None], self.d[None])
        nl = round3(np.clip(self.l[None] + self.o.dt * f["dl"], 0, 1))[0]
        nd = round3(np.clip(self.d[None] + self.o.dt * f["dd"], 0, 1))[0]
        self.l[1:-1], self.d[1:-1] = nl[1:-1], nd[1:-1]
        self.stepmax[2 * self.j] = int(np.rint(self.l[1:-1].max() * 1000))
        self.stepmax[2 * self.j + 1] = int(np.rint(self.d[1:-1].max() * 1000))
        self.j += 1
        # update_L (reference :463-473), ramp_up_down off
        self.clk.step_count += 1
        self.clk.L = max(min(self.clk.L + self.clk.dL, self.clk.max_L), self.clk.min_L)

    def halo_wrap(self):
        self.l[0], self.l[-1] = self.l[-2].copy(), self.l[1].copy()
        self.d[0], self.d[-1] = self.d[-2].copy(), self.d[1].copy()

    def _absorb_halo(self):
        """Unpack the ghost rows received by the last exchange (packed milli-cover words like the device rows)."""
        if getattr(self, "_halo", None) is not None:
            for name, row in (("recv_top", 0), ("recv_bottom", self.rows + 1)):
                w = self._halo[name].numpy().astype(np.int64)
                self.l[row] = (w & 0xffff) / 1000.0
                self.d[row] = (w >> 16) / 1000.0
            self._halo = None

    def exch_tensor(self, gain=True, act=True):
        return self.exch[(0 if gain else self.n):(2 * self.n if act else self.n)]

    def stepmax_tensor(self, K):
        return self.stepmax[:2 * K]

    def halo_tensors(self):
        t = self._torch

        def pack(r):
            kl = np.rint(self.l[r] * 1000).astype(np.int64)
            kd = np.rint(self.d[r] * 1000).astype(np.int64)
            return t.from_numpy((kl | (kd << 16)).astype(np.int32))

        self._halo = {"send_top": pack(1), "send_bottom": pack(self.rows), "recv_top": t.zeros(self.N, dtype=t.int32),
                      "recv_bottom": t.zeros(self.N, dtype=t.int32)}
        return self._halo["send_top"], self._halo["send_bottom"], self._halo["recv_top"], self._halo["recv_bottom"]

    def end_chunk(self, K):
        m = self.stepmax[:2 * K].numpy().reshape(K, 2).max(axis=1)
        first = -1
        for j in range(K):
            if m[j] > 5:
                self.done_at += 1
            elif first < 0:
                first = j
        self.stepmax[:2 * self.j] = 0
        self.j = 0
        return first

    def reset_lifespans(self):
        self.done_at = 0
        if hasattr(self, "ada"):
            self.ada[:] = 0

    def lifespans(self):
        return self.done_at, self.ada.copy()

    def agents(self):
        return self.xy.copy(), self.st.copy()

    def covers(self):
        self._absorb_halo()
        return np.stack([self.l[1:-1], self.d[1:-1]])
