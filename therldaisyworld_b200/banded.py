"""A single giant toroidal RLDaisyWorld (BASELINE config 5), row-banded over the ranks of one node.

The world is ONE batch element of the reference environment (``daisy/daisy_world_rl.py``, ``batch_size == 1``) whose
grid does not fit in shared memory: rank r owns rows ``[r*N/R, (r+1)*N/R)`` on its GPU (a ``dwt_handle`` of
``include/daisyworld_b200_tiled.h``) plus a ghost row above and below.  Agents are replicated on every rank.

One env step = the phase sequence of the header, with exactly these exchanges between ranks (``torch.distributed``:
NCCL over NVLink on GPUs, gloo in the CPU tests -- plumbing only, the data path is the CUDA kernels):

  * ONE SUM all-reduce per step over ``[gain(j-1) | act(j)]``: only the owner band of a cell knows what was eaten there,
    only the owner band of an agent sees its neighbourhood (greedy / anti-greedy).  Finishing step j-1 (state += gain,
    clip, reward/done) is deferred until the decisions of step j exist, so both vectors travel together;
  * after the edge tile rows of the stencil: the band's first/last row goes to the upper/lower neighbour's ghost row
    (ring, toroidal) on a side stream while the interior tile rows are still being computed;
  * per chunk of steps: MAX all-reduce of the per-step cover maxima (lifespan bookkeeping, ``grid_done``).

``world_size == 1`` needs none of them (``dwt_halo_wrap`` closes the torus locally).

``mode="p2p"`` (GPUs of one node) takes the collectives off the step path altogether: the ranks map each other's
exchange vectors, barrier flags and lattice buffers through CUDA IPC once, and a step is ``dwt_step_p2p`` -- the owner of
an agent / the winner of a graze stores straight into every rank's exchange vector over NVLink, edge rows are pushed
into the neighbours' ghost rows, ranks meet at device-side flag barriers.  Only the per-chunk MAX all-reduce remains.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import DwClock, DwtPtrs, DW_POLICY
from .env import make_clock_struct, make_config_struct, set_default_attributes, set_default_kernels

_WORLD_POLICIES = ("greedy", "antigreedy", "eps_greedy")      # decisions that (may) read the grid: need the act exchange


def band_rows(N, world_size, rank):
    """Rows [lo, hi) of rank's band; N must split into equal multiples of 64."""
    if N % (64 * world_size):
        raise ValueError(f"N={N} must be a multiple of 64*world_size={64 * world_size}")
    R = N // world_size
    return rank * R, (rank + 1) * R


class DistComm:
    """The three exchanges of a banded step over torch.distributed (NCCL on GPUs, gloo on CPU)."""

    def __init__(self, rank, world_size, group=None):
        import torch.distributed as dist
        self.dist, self.rank, self.world_size, self.group = dist, int(rank), int(world_size), group

    def all_reduce_sum(self, t):
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)

    def all_reduce_max(self, t):
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX, group=self.group)

    def exchange_halos(self, send_top, send_bottom, recv_top, recv_bottom):
        dist = self.dist
        up, down = (self.rank - 1) % self.world_size, (self.rank + 1) % self.world_size
        if self.group is not None:
            up, down = dist.get_global_rank(self.group, up), dist.get_global_rank(self.group, down)
        # tag 0: downward message (my last row -> lower neighbour's top ghost); tag 1: upward. Posting order matters for
        # NCCL when up == down (two ranks): the first send pairs with the peer's first receive.
        ops = [dist.P2POp(dist.isend, send_bottom, down, group=self.group, tag=0),
               dist.P2POp(dist.isend, send_top, up, group=self.group, tag=1),
               dist.P2POp(dist.irecv, recv_top, up, group=self.group, tag=0),
               dist.P2POp(dist.irecv, recv_bottom, down, group=self.group, tag=1)]
        for req in dist.batch_isend_irecv(ops):
            req.wait()


class _DevArray:
    """Minimal __cuda_array_interface__ carrier so torch can view library-owned device memory without a copy."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


class DeviceBand:
    """One rank's band on its GPU (ctypes over the dwt_* C-ABI)."""

    def __init__(self, params, N, n_agents, row0, rows, n_ranks, device=0):
        import torch
        self._torch = torch
        self._lib = _lib.load()
        self.N, self.n, self.row0, self.rows, self.n_ranks, self.device = int(N), int(n_agents), int(row0), int(rows), int(n_ranks), int(device)
        cfg = make_config_struct(params, 1, N, n_agents, device)
        h = C.c_void_p()
        rc = self._lib.dwt_create(C.byref(cfg), self.rows, self.row0, self.n_ranks, C.byref(h))
        if rc != 0:
            msg = self._lib.dwt_last_error(None)
            raise _lib.DaisyWorldError(f"dwt_create failed (code {rc}): {msg.decode() if msg else ''}")
        self._h = h
        self._views = {}
        self._side = None

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._lib.dwt_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def _check(self, rc, what):
        _lib.check_tiled(self._lib, self._h, rc, what)

    # ---- state
    def set_params(self, params):
        cfg = make_config_struct(params, 1, self.N, self.n, self.device)
        self._check(self._lib.dwt_set_config(self._h, C.byref(cfg)), "dwt_set_config")

    def set_clock(self, clk):
        self._check(self._lib.dwt_set_clock(self._h, C.byref(clk)), "dwt_set_clock")

    def set_epsilon(self, epsilon):
        self._check(self._lib.dwt_set_epsilon(self._h, float(epsilon)), "dwt_set_epsilon")

    def get_clock(self):
        clk = DwClock()
        self._check(self._lib.dwt_get_clock(self._h, C.byref(clk)), "dwt_get_clock")
        return clk

    def upload(self, light_rows, dark_rows, agent_indices, agent_states):
        """light_rows/dark_rows: [rows + 2, N] (band with its ghost rows); agents: all n, global coordinates."""
        pd = C.POINTER(C.c_double)
        l = np.ascontiguousarray(light_rows, dtype=np.float64)
        d = np.ascontiguousarray(dark_rows, dtype=np.float64)
        assert l.shape == d.shape == (self.rows + 2, self.N)
        self._check(self._lib.dwt_upload_covers(self._h, l.ctypes.data_as(pd), d.ctypes.data_as(pd)), "dwt_upload_covers")
        if self.n:
            ai = np.ascontiguousarray(agent_indices, dtype=np.int64).reshape(self.n, 2)
            st = np.ascontiguousarray(agent_states, dtype=np.float64).reshape(self.n)
            self._check(self._lib.dwt_upload_agents(self._h, ai.ctypes.data_as(C.POINTER(C.c_int64)), st.ctypes.data_as(pd)),
                        "dwt_upload_agents")
        self._views = {}

    def init_random(self, seed, params):
        self._check(self._lib.dwt_init_random(self._h, C.c_uint64(seed), params.light_proportion, params.dark_proportion,
                                              params.initial_al, params.initial_ad), "dwt_init_random")
        self._views = {}

    # ---- phases
    def decide(self, policy, actions_step=None, seed=0):
        a = None
        if policy == "replay":
            a = np.ascontiguousarray(np.asarray(actions_step).reshape(self.n), dtype=np.int8)
        self._check(self._lib.dwt_decide(self._h, DW_POLICY[policy], None if a is None else a.ctypes.data_as(C.POINTER(C.c_int8)),
                                         C.c_uint64(seed)), "dwt_decide")

    def move_graze(self):
        self._check(self._lib.dwt_move_graze(self._h), "dwt_move_graze")

    def finish_agents(self):
        self._check(self._lib.dwt_finish_agents(self._h), "dwt_finish_agents")

    def stencil(self, part=0):
        self._check(self._lib.dwt_stencil(self._h, int(part)), "dwt_stencil")

    def halo_wrap(self):
        self._check(self._lib.dwt_halo_wrap(self._h), "dwt_halo_wrap")

    def stencil_and_exchange(self, comm):
        """Edge tile rows, then the halo exchange on a side stream overlapped with the interior tile rows."""
        torch = self._torch
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device)
            self._edge_done = torch.cuda.Event()
        self.stencil(1)
        main = torch.cuda.current_stream(self.device)
        self._edge_done.record(main)
        with torch.cuda.stream(self._side):
            self._side.wait_event(self._edge_done)
            comm.exchange_halos(*self.halo_tensors())
        self.stencil(2)
        main.wait_stream(self._side)

    # ---- peer-memory mode
    def set_stream(self, cuda_stream):
        self._check(self._lib.dwt_set_stream(self._h, C.c_void_p(int(cuda_stream))), "dwt_set_stream")

    def ipc_export(self):
        buf = C.create_string_buffer(4 * 64)
        self._check(self._lib.dwt_ipc_export(self._h, buf), "dwt_ipc_export")
        return bytes(buf.raw)

    def ipc_attach(self, rank, blobs):
        blob = b"".join(blobs)
        self._check(self._lib.dwt_ipc_attach(self._h, int(rank), len(blobs), C.create_string_buffer(blob, len(blob))), "dwt_ipc_attach")

    def peer_buffers(self):
        out = (C.c_void_p * 4)()
        self._check(self._lib.dwt_get_peer_buffers(self._h, out), "dwt_get_peer_buffers")
        return [int(v or 0) for v in out]

    def attach_peers(self, rank, tables):
        """Same-process peers: tables[r] = peer_buffers() of rank r."""
        flat = (C.c_void_p * (4 * len(tables)))(*[p for t in tables for p in t])
        self._check(self._lib.dwt_attach_peers(self._h, int(rank), len(tables), flat), "dwt_attach_peers")

    def step_p2p(self, policy, actions_step=None, seed=0):
        a = None
        if policy == "replay":
            a = np.ascontiguousarray(np.asarray(actions_step).reshape(self.n), dtype=np.int8)
        self._check(self._lib.dwt_step_p2p(self._h, DW_POLICY[policy], None if a is None else a.ctypes.data_as(C.POINTER(C.c_int8)),
                                           C.c_uint64(seed)), "dwt_step_p2p")

    def flush_p2p(self):
        self._check(self._lib.dwt_flush_p2p(self._h), "dwt_flush_p2p")

    def cover_checksum(self):
        out = (C.c_uint64 * 4)()
        self._check(self._lib.dwt_cover_checksum(self._h, out), "dwt_cover_checksum")
        return [int(v) for v in out]

    def time_stencil(self, reps=20):
        us = C.c_double(0.0)
        self._check(self._lib.dwt_debug_time_stencil(self._h, int(reps), C.byref(us)), "dwt_debug_time_stencil")
        return us.value

    def peer_timed_out(self):
        v = C.c_int32(0)
        self._check(self._lib.dwt_peer_status(self._h, C.byref(v)), "dwt_peer_status")
        return bool(v.value)

    def run_local(self, K, policy, actions=None, seed=0):
        """K steps without any exchange (a band that is the whole torus)."""
        a = None
        if policy == "replay":
            a = np.ascontiguousarray(np.asarray(actions).reshape(-1, self.n)[:K], dtype=np.int8)
            assert a.shape[0] == K
        self._check(self._lib.dwt_run(self._h, int(K), DW_POLICY[policy], None if a is None else a.ctypes.data_as(C.POINTER(C.c_int8)),
                                      C.c_uint64(seed)), "dwt_run")

    def end_chunk(self, K):
        first = C.c_int32(-1)
        self._check(self._lib.dwt_end_chunk(self._h, int(K), C.byref(first)), "dwt_end_chunk")
        return int(first.value)

    # ---- exchange buffers as torch CUDA tensors (zero-copy views of library memory)
    def _ptrs(self):
        p = DwtPtrs()
        self._check(self._lib.dwt_get_ptrs(self._h, C.byref(p)), "dwt_get_ptrs")
        return p

    def _view(self, ptr, shape, typestr):
        key = (int(ptr), tuple(shape), typestr)
        t = self._views.get(key)
        if t is None:
            t = self._torch.as_tensor(_DevArray(ptr, shape, typestr), device=f"cuda:{self.device}")
            self._views[key] = t
        return t

    def exch_tensor(self, gain=True, act=True):
        """The exchange vector [gain[n] | act[n]] (or one half of it) as a torch view."""
        t = self._view(self._ptrs().gain, (2 * max(self.n, 1),), "<f8")
        return t[(0 if gain else self.n):(2 * self.n if act else self.n)]

    def stepmax_tensor(self, K):
        return self._view(self._ptrs().stepmax, (4096 * 2,), "<i4")[:2 * K]

    def halo_tensors(self):
        """(send_top, send_bottom, recv_top, recv_bottom): whole stored rows (N + 8 words, ghost columns included) of the
        lattice buffer the current step writes, as int32 views."""
        p = self._ptrs()
        return tuple(self._view(q, (self.N + 8,), "<i4") for q in (p.send_top, p.send_bottom, p.recv_top, p.recv_bottom))

    # ---- getters
    def reset_lifespans(self):
        self._check(self._lib.dwt_reset_lifespans(self._h), "dwt_reset_lifespans")

    def lifespans(self):
        done_at = C.c_int64(0)
        ada = np.zeros((self.n,), dtype=np.int64)
        self._check(self._lib.dwt_get_lifespans(self._h, C.byref(done_at), ada.ctypes.data_as(C.POINTER(C.c_int64))), "dwt_get_lifespans")
        return int(done_at.value), ada

    def agents(self):
        ai = np.zeros((self.n, 2), dtype=np.int64)
        st = np.zeros((self.n,), dtype=np.float64)
        if self.n:
            self._check(self._lib.dwt_get_agents(self._h, ai.ctypes.data_as(C.POINTER(C.c_int64)), st.ctypes.data_as(C.POINTER(C.c_double))),
                        "dwt_get_agents")
        return ai, st

    def reward_done(self):
        r = np.zeros((self.n,), dtype=np.float64)
        d = np.zeros((self.n,), dtype=np.uint8)
        if self.n:
            self._check(self._lib.dwt_get_reward_done(self._h, r.ctypes.data_as(C.POINTER(C.c_double)), d.ctypes.data_as(C.POINTER(C.c_uint8))),
                        "dwt_get_reward_done")
        return r, d.astype(bool)

    def covers(self):
        out = np.empty((2, self.rows, self.N))
        pd = C.POINTER(C.c_double)
        self._check(self._lib.dwt_get_covers(self._h, out[0].ctypes.data_as(pd), out[1].ctypes.data_as(pd)), "dwt_get_covers")
        return out

    def grid(self):
        out = np.empty((7, self.rows, self.N))
        self._check(self._lib.dwt_get_grid(self._h, out.ctypes.data_as(C.POINTER(C.c_double))), "dwt_get_grid")
        return out

    def slow_count(self):
        c = C.c_uint64(0)
        self._check(self._lib.dwt_debug_slow_count(self._h, C.byref(c)), "dwt_debug_slow_count")
        return int(c.value)

    def synchronize(self):
        self._check(self._lib.dwt_synchronize(self._h), "dwt_synchronize")


class BandedDaisyWorld:
    """Host-side driver of one giant world over `world_size` bands.

    Carries the reference's public constants as attributes (same names as RLDaisyWorld; mutate them before reset()).
    `band_factory(params, N, n_agents, row0, rows, n_ranks)` builds this rank's band: DeviceBand on a GPU (default), an
    oracle-backed stand-in in the CPU tests."""

    def __init__(self, grid_dimension, n_agents, rank=0, world_size=1, group=None, device=0, band_factory=None, comm=None,
                 mode="nccl", **kwargs):
        set_default_attributes(self, grid_dimension=grid_dimension, n_agents=n_agents, **kwargs)
        set_default_kernels(self)
        self.batch_size = 1
        self.rank, self.world_size, self.device = int(rank), int(world_size), int(device)
        self.comm = comm if (comm is not None or self.world_size == 1) else DistComm(rank, world_size, group)
        self.row0, hi = band_rows(self.dim, self.world_size, self.rank)
        self.rows = hi - self.row0
        factory = band_factory or (lambda p, N, n, r0, R, nr: DeviceBand(p, N, n, r0, R, nr, device=self.device))
        self.band = factory(self, self.dim, self.n_agents, self.row0, self.rows, self.world_size)
        self.L = self.min_L
        self.dL = (self.max_L - self.min_L) / self.ramp_period
        self.step_count = 0
        self._pending = 0          # steps recorded since the last end_chunk
        self._gain_pending = False # multi-rank: gains of the last step not yet summed / applied
        self.first_done_step = None
        self.mode = mode if self.world_size > 1 else "local"
        if self.mode == "p2p" and comm is None:
            # one-off exchange of the CUDA IPC handles (any transport would do; torch.distributed is at hand)
            blobs = [None] * self.world_size
            self.comm.dist.all_gather_object(blobs, self.band.ipc_export(), group=group)
            self.band.ipc_attach(self.rank, blobs)

    # ---- reset
    def _reset_clock(self):
        self.L = self.min_L
        self.dL = (self.max_L - self.min_L) / self.ramp_period
        self.step_count = 0
        self.band.set_params(self)
        self.band.set_clock(make_clock_struct(self))
        self.band.reset_lifespans()
        self._pending = 0
        self._gain_pending = False
        self.first_done_step = None

    def load_state(self, light, dark, agent_indices, agent_states):
        """reset() from explicit GLOBAL arrays: light/dark [N,N] (may be off the 0.001 lattice), agents [n,2], [n]."""
        self._reset_clock()
        N = self.dim
        rows = np.arange(self.row0 - 1, self.row0 + self.rows + 1) % N
        self.band.upload(np.asarray(light)[rows], np.asarray(dark)[rows], agent_indices, agent_states)

    def reset_on_device(self, seed=0):
        """reset() with the state drawn on the device (same distribution as the reference, counter RNG keyed by the
        global cell index: identical for every banding)."""
        self._reset_clock()
        self.band.init_random(seed, self)

    # ---- stepping
    def step(self, policy="greedy", actions_step=None, seed=0):
        b, comm = self.band, self.comm
        if self.mode == "p2p":
            b.step_p2p(policy, actions_step, seed)
            self._gain_pending = bool(self.n_agents)
            self._pending += 1
            return
        b.decide(policy, actions_step, seed)
        if comm is None:
            b.move_graze()
            b.finish_agents()
            b.stencil(0)
            b.halo_wrap()
        else:
            need_act = bool(self.n_agents) and policy in _WORLD_POLICIES
            if self._gain_pending or need_act:
                comm.all_reduce_sum(b.exch_tensor(gain=self._gain_pending, act=need_act))
            if self._gain_pending:
                b.finish_agents()                    # step j-1, before its gz flags are overwritten
            b.move_graze()
            self._gain_pending = bool(self.n_agents)
            if hasattr(b, "stencil_and_exchange"):
                b.stencil_and_exchange(comm)
            else:
                b.stencil(0)
                comm.exchange_halos(*b.halo_tensors())
        self._pending += 1

    def _check_peers(self):
        """Peer-memory mode: a device-side flag barrier gives up after a bounded spin (csrc/dw_tiled.cuh, ~4 s or
        DW_PEER_TIMEOUT_MS) and the step then continues on stale ghost rows / exchange vectors. Every point where results
        leave the device goes through here, so a stalled peer raises instead of yielding silently wrong grids and lifespans."""
        if self.mode == "p2p" and self.band.peer_timed_out():
            raise _lib.DaisyWorldError(f"rank {self.rank}: a peer-memory barrier timed out (a peer rank stalled or died); the state of "
                                  "this world is not trustworthy -- reset() before continuing")

    def _flush(self):
        """Finish the last step's agents (multi-rank: the deferred gain all-reduce)."""
        if self._gain_pending:
            if self.mode == "p2p":
                self.band.flush_p2p()
                self._check_peers()
            else:
                self.comm.all_reduce_sum(self.band.exch_tensor(gain=True, act=False))
                self.band.finish_agents()
            self._gain_pending = False

    def end_chunk(self):
        """Fold the per-step cover maxima of the steps since the last call into the lifespan counter."""
        self._flush()
        K = self._pending
        if K == 0:
            return
        if self.comm is not None:
            self.comm.all_reduce_max(self.band.stepmax_tensor(K))
        first = self.band.end_chunk(K)
        self._check_peers()
        if first >= 0 and self.first_done_step is None:
            self.first_done_step = self.step_count + first + 1      # 1-based count of steps at which grid_done first held
        self.step_count += K
        self._pending = 0
        clk = self.band.get_clock()
        self.L, self.dL, self.min_L, self.max_L = clk.L, clk.dL, clk.min_L, clk.max_L

    def run(self, K, policy="greedy", actions=None, seed=0, chunk=256):
        """K env steps. actions: [K, n] ints 0..8 for policy='replay'. Lifespan counters are folded every `chunk` steps."""
        if self.world_size == 1 and hasattr(self.band, "run_local"):
            done = 0
            while done < K:
                k = min(chunk, K - done)
                self.band.run_local(k, policy, None if actions is None else np.asarray(actions)[done:done + k], seed)
                self._pending += k
                self.end_chunk()
                done += k
            return K
        for j in range(K):
            self.step(policy, None if actions is None else np.asarray(actions)[j], seed)
            if self._pending >= chunk:
                self.end_chunk()
        self.end_chunk()
        return K

    # ---- results
    def lifespans(self):
        """(done_at, agents_done_at[n]) -- the notebook counters (greedy_longevity_abatement.ipynb cell 2)."""
        self.end_chunk()
        return self.band.lifespans()

    def agents(self):
        self._flush()
        return self.band.agents()

    def cover_checksum(self):
        """Exact position-weighted checksum of the WHOLE world's covers (this rank's part summed over the ranks mod 2^64):
        identical for every banding of the same world (dwt_cover_checksum)."""
        self._flush()
        cs = self.band.cover_checksum()
        if self.comm is not None and hasattr(self.comm, "dist"):
            # int64 all-reduce wraps mod 2^64 like the device sums; the four words travel as signed values
            import torch
            t = torch.tensor([c - (1 << 64) if c >= (1 << 63) else c for c in cs], dtype=torch.int64, device=f"cuda:{self.device}")
            self.comm.dist.all_reduce(t, group=self.comm.group)
            cs = [int(v) & ((1 << 64) - 1) for v in t.tolist()]
        return cs

    def local_covers(self):
        """[2, rows, N]: light and dark cover of this rank's band."""
        return self.band.covers()

    def local_grid(self):
        """[7, rows, N]: env.grid[0] restricted to this rank's rows."""
        self._flush()
        return self.band.grid()
