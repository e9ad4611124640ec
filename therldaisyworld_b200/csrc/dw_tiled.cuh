// Single large world ("giant grid", BASELINE config 5): one toroidal N x N world too large for shared memory, held as a
// row band of `R` rows per rank (R = N on one GPU) on the packed 0.001 lattice, padded with ghost cells:
//
//     lat[(R + 2) x pitch],  pitch = N + 8 words
//       row 0        ghost: the row above the band (toroidal wrap, or the neighbouring rank's last row)
//       rows 1..R    the band
//       row R + 1    ghost: the row below the band
//       word 3 / words 4..N+3 / word N+4 of a row: ghost of column N-1 / the N columns / ghost of column 0
//
// With the ghosts in place every 64x64 output tile needs one in-bounds 66-row x 72-word box of the source, which one
// elected thread fetches with a single TMA (cp.async.bulk.tensor.2d) into shared memory; 256 threads then run the same
// 4x4-cells-per-thread fast path as the ensemble kernel (dw_tile_core) and write their rows straight back to HBM.
// HBM traffic is 8 B per cell-update (4 B in, 4 B out); at the FP64-bound rate that is ~2 TB/s of the 6.5 TB/s measured,
// so this kernel is bound by the FP64 pipe like the SMEM-resident one.
//
// Agents of a giant world are many (thousands) and are replicated on every rank; the reference's sequential loop
// (daisy_world_rl.py:186-216) is resolved in parallel: decisions from the pre-move state, moves independent, and the
// grazers of one cell are ordered by an atomicMin claim (claim[(R+2) x N], INT_MAX when idle) on the agent index (the lowest index eats, later ones find the
// cell empty -- SURVEY App. B.8).  A rank applies the grazes that land in its band; grazes landing in its ghost rows only
// zero the ghost copy.  Per-agent food gains are summed over ranks by the caller (one all-reduce; exactly one rank
// contributes a non-zero term).
#pragma once
#include <cuda.h>

#include "dw_fused.cuh"

struct BandGeom {
    int N;       // global grid side
    int R;       // rows of this band
    int row0;    // global index of the band's first row
    int pitch;   // words (or doubles) per stored row
    int c0;      // index of column 0 inside a stored row (4 for the lattice, 0 for the fp64 planes of the first step)
};

// local stored-row indices (0..R+1) at which global row x is held by this band: interior and/or ghost copies
template <class F>
__device__ __forceinline__ void band_row_images(const BandGeom &G, int x, F f) {
    int rel = x - G.row0;
    rel = rel < 0 ? rel + G.N : rel;
    if (rel < G.R) f(rel + 1);
    if (rel == G.N - 1) f(0);
    if (rel == (G.R == G.N ? 0 : G.R)) f(G.R + 1);
}
template <class F>
__device__ __forceinline__ void band_col_images(const BandGeom &G, int y, F f) {
    f(G.c0 + y);
    if (G.c0 > 0) {
        if (y == G.N - 1) f(G.c0 - 1);
        if (y == 0) f(G.c0 + G.N);
    }
}
// interior stored row of global row x, or -1 if this band does not own it
__device__ __forceinline__ int band_owned_row(const BandGeom &G, int x) {
    int rel = x - G.row0;
    rel = rel < 0 ? rel + G.N : rel;
    return rel < G.R ? rel + 1 : -1;
}

// ---- cell accessors: packed lattice (steady state) or fp64 cover planes (the off-lattice state right after reset) ----
struct LatCells {
    uint32_t *p;
    __device__ __forceinline__ double food(size_t i) const { return dw_food(p[i]); }
    __device__ __forceinline__ void zero(size_t i) const { p[i] = 0u; }
};
struct PlaneCells {
    double *l, *d;
    __device__ __forceinline__ double food(size_t i) const { return l[i] + d[i]; }
    __device__ __forceinline__ void zero(size_t i) const { l[i] = 0.0; d[i] = 0.0; }
};
// column index with toroidal wrap for planes without column ghosts
__device__ __forceinline__ int band_col(const BandGeom &G, int y) {
    if (G.c0 > 0) return G.c0 + y;                       // ghosts at c0-1 and c0+N make y = -1 and y = N valid
    return y < 0 ? y + G.N : (y >= G.N ? y - G.N : y);
}

// ---- peer-memory mode (one process per GPU, P2P stores over NVLink / NVSwitch, no NCCL on the step path) --------------
// Every rank maps every other rank's exchange vector, flag array and lattice buffers (CUDA IPC).  The owner of an agent /
// the winner of a graze stores its result straight into ALL ranks' exchange vectors (exactly one writer per entry, so
// nothing has to be reduced), a band pushes its edge rows into the neighbours' ghost rows, and ranks meet at flag
// barriers (k_peer_barrier) instead of collectives.
#define DWT_MAX_RANKS 8
struct PeerTable {
    double *exch[DWT_MAX_RANKS];           // [gain1 | gain0 | act], n doubles each
    unsigned int *flags[DWT_MAX_RANKS];    // [DWT_MAX_RANKS] barrier epochs, slot r written by rank r
    int rank, R, on;                       // on == 0: single-process / NCCL mode (local stores only)
    int pad_;
    long long timeout_clocks;              // bound of a barrier spin (default 8e9 ~ 4 s; DW_PEER_TIMEOUT_MS); <= 0: default
};

// All ranks meet here: every prior write of this rank (to its own or to peer memory) is visible to a peer that has seen
// the flag. Spins are bounded (~seconds) so that a missing peer produces an error flag instead of a hung GPU.
__device__ __forceinline__ void dw_peer_barrier_body(const PeerTable &PT, unsigned int epoch, unsigned int *timed_out, int p) {
    if (p >= PT.R) return;
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(PT.flags[p] + PT.rank), "r"(epoch) : "memory");
    const unsigned int *mine = PT.flags[PT.rank] + p;
    const long long t0 = clock64(), limit = PT.timeout_clocks > 0 ? PT.timeout_clocks : 8000000000ll;
    for (;;) {
        unsigned int v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
        if ((int)(v - epoch) >= 0) break;
        if (clock64() - t0 > limit) { *timed_out = 1u; break; }
        __nanosleep(100);
    }
    __threadfence_system();
}
__global__ void k_peer_barrier(PeerTable PT, unsigned int epoch, unsigned int *timed_out) {
    dw_peer_barrier_body(PT, epoch, timed_out, threadIdx.x);
}
// Tail of a multi-block kernel that ends in a barrier: every block calls this after its own (peer) stores; the block
// that arrives last at the ticket counter runs the barrier, so the flag is raised only after ALL blocks' stores.
__device__ __forceinline__ void dw_last_block_barrier(const PeerTable &PT, unsigned int epoch, unsigned int *timed_out, unsigned int *ticket) {
    __shared__ int s_last;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(ticket, 1u);
        s_last = (t == gridDim.x - 1);
        if (s_last) *ticket = 0u;                        // re-armed for the next launch (stream order)
        __threadfence();
    }
    __syncthreads();
    if (s_last) dw_peer_barrier_body(PT, epoch, timed_out, threadIdx.x);
}

// a band's first / last stored row (ghost columns included) -> the neighbours' ghost rows
// ... and then the closing barrier of the step, raised by the block that finishes last.
__global__ void __launch_bounds__(256) k_band_push_halo(const uint32_t *__restrict__ lat, int R, int pitch, uint32_t *up_ghost_bottom,
                                                        uint32_t *down_ghost_top, PeerTable PT, unsigned int epoch, unsigned int *timed_out,
                                                        unsigned int *ticket) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < pitch) {
        up_ghost_bottom[c] = lat[(size_t)pitch + c];
        down_ghost_top[c] = lat[(size_t)R * pitch + c];
    }
    dw_last_block_barrier(PT, epoch, timed_out, ticket);
}

// ---- agents --------------------------------------------------------------------------------------------------------
template <class Cells>
__device__ __forceinline__ void k_band_decide_one(const BandGeom &G, const Cells &C, const int32_t *__restrict__ xy, int n, int policy,
                                                  const int8_t *__restrict__ replay, uint64_t seed, uint32_t step, const PeerTable &PT,
                                                  double *__restrict__ act, int i) {
    const int x = xy[2 * i], y = xy[2 * i + 1];
    const int lr = band_owned_row(G, x);
    double out = 0.0;
    if (policy == DW_POLICY_REPLAY || policy == DW_POLICY_NONE || policy == DW_POLICY_RANDOM) {
        // world-independent: every rank computes the same action for every agent (no exchange needed)
        int a = 0;
        if (policy == DW_POLICY_REPLAY) a = replay[i];
        else if (policy == DW_POLICY_RANDOM) a = (int)(dw_hash_rng(seed, 0u, (uint32_t)i, step) % 9u);
        out = (double)(a + 1);
    } else if (lr >= 0) {
        const size_t rowm = (size_t)(lr - 1) * G.pitch, row = (size_t)lr * G.pitch, rowp = (size_t)(lr + 1) * G.pitch;
        const double food[4] = {C.food(row + band_col(G, y - 1)), C.food(rowm + band_col(G, y)), C.food(rowp + band_col(G, y)),
                                C.food(row + band_col(G, y + 1))};
        out = (double)(dw_greedy_pick(food, policy == DW_POLICY_GREEDY) + 1);
        if (PT.on) {                                     // owner publishes the decision to every rank (act = exch + 2n)
            for (int p = 0; p < PT.R; ++p) PT.exch[p][2 * (size_t)n + i] = out;
            return;
        }
    } else if (PT.on) return;                            // peer-memory mode: the owner rank writes this entry
    act[i] = out;
}

// Phase 1: the owner band of each agent decides its action from the pre-move state. act[i] = action + 1 for owned
// agents, 0 otherwise (the caller sums act over ranks). Policies that do not look at the world (replay, none, random)
// fill every entry on every rank and need no exchange.
template <class Cells>
__global__ void __launch_bounds__(256) k_band_decide(BandGeom G, Cells C, const int32_t *__restrict__ xy, int n, int policy,
                                                     const int8_t *__restrict__ replay, uint64_t seed, uint32_t step, PeerTable PT,
                                                     double *__restrict__ act, unsigned int epoch, unsigned int *timed_out,
                                                     unsigned int *ticket) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) k_band_decide_one<Cells>(G, C, xy, n, policy, replay, seed, step, PT, act, i);
    if (epoch) dw_last_block_barrier(PT, epoch, timed_out, ticket);     // peer-memory mode, world-reading policies
}

// Phase 2 (replicated on every rank): pay agent_gamma, move the living, and file the graze claims of cells this band owns.
// gz[i] = 1 if agent i grazes this step.
__global__ void __launch_bounds__(256) k_band_move_claim(BandGeom G, double agent_gamma, int32_t *xy, double *st, int n,
                                                         const double *__restrict__ act, int *claim, uint8_t *gz) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int a = (int)act[i] - 1;
    const double s = st[i] - agent_gamma;
    st[i] = s;
    int x = xy[2 * i], y = xy[2 * i + 1];
    uint8_t g = 0;
    if (s > 0.0) {
        if (a != 8) {
            const int d = (a & 2) ? 1 : -1;              // a & 3 = 0: y-1, 1: x-1, 2: x+1, 3: y+1
            if (((a + 1) & 2) == 0) { y += d; y = y < 0 ? y + G.N : (y >= G.N ? y - G.N : y); }
            else { x += d; x = x < 0 ? x + G.N : (x >= G.N ? x - G.N : x); }
            xy[2 * i] = x;
            xy[2 * i + 1] = y;
        }
        if (a > 4) {
            g = 1;
            const int lr = band_owned_row(G, x);
            if (lr >= 0) atomicMin(claim + (size_t)lr * G.N + y, i);
        }
    }
    gz[i] = g;
}

// Peer-memory mode: phase 4 of step j-1 and phase 2 of step j in one launch (the finish is deferred until the gains are
// known to be complete, i.e. after the closing barrier of step j-1). The two steps use different claim arrays (step
// parity), so returning step j-1's claims to "idle" cannot collide with step j's atomicMin on the same cell.
__global__ void __launch_bounds__(256) k_band_finish_move_claim(BandGeom G, double agent_gamma, int32_t *xy, double *st, int n,
                                                                const double *__restrict__ act, int *claim_new, uint8_t *gz, int do_finish,
                                                                double *gain_prev, int *claim_prev, double *reward, uint8_t *done,
                                                                int64_t *agents_done_at) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = st[i];
    int x = xy[2 * i], y = xy[2 * i + 1];
    if (do_finish) {
        const double g = gain_prev[i];
        gain_prev[i] = 0.0;
        if (gz[i]) {
            s = s + g;
            const int lr = band_owned_row(G, x);
            if (lr >= 0) claim_prev[(size_t)lr * G.N + y] = 0x7fffffff;
        }
        s = dw_clip01(s);
        reward[i] = s;
        done[i] = s < 0.1;
        agents_done_at[i] += (s < 0.1) ? 0 : 1;
    }
    const int a = (int)act[i] - 1;
    s = s - agent_gamma;
    st[i] = s;
    uint8_t g = 0;
    if (s > 0.0) {
        if (a != 8) {
            const int d = (a & 2) ? 1 : -1;
            if (((a + 1) & 2) == 0) { y += d; y = y < 0 ? y + G.N : (y >= G.N ? y - G.N : y); }
            else { x += d; x = x < 0 ? x + G.N : (x >= G.N ? x - G.N : x); }
            xy[2 * i] = x;
            xy[2 * i + 1] = y;
        }
        if (a > 4) {
            g = 1;
            const int lr = band_owned_row(G, x);
            if (lr >= 0) atomicMin(claim_new + (size_t)lr * G.N + y, i);
        }
    }
    gz[i] = g;
}

// Phase 3: the claim winner of an owned cell eats it and clears every stored copy; grazes that land in a ghost row only
// clear the ghost copy (the owner rank does the eating). gain[i] = food eaten by agent i on THIS rank (0 elsewhere).
template <class Cells>
__global__ void __launch_bounds__(256) k_band_graze(BandGeom G, Cells C, const int32_t *__restrict__ xy, int n,
                                                    const uint8_t *__restrict__ gz, const int *__restrict__ claim, double *__restrict__ gain,
                                                    PeerTable PT, int gain_off) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double got = 0.0;
    bool won = false;
    if (gz[i]) {
        const int x = xy[2 * i], y = xy[2 * i + 1];
        const int lr = band_owned_row(G, x);
        bool clear = true;
        if (lr >= 0) {
            if (claim[(size_t)lr * G.N + y] == i) { got = C.food((size_t)lr * G.pitch + G.c0 + y); won = true; }
            else clear = false;                          // a lower-index agent eats here; it also does the clearing
        }
        if (clear)
            band_row_images(G, x, [&](int r) { band_col_images(G, y, [&](int cc) { C.zero((size_t)r * G.pitch + cc); }); });
    }
    if (PT.on) {                                         // the winner publishes to every rank; the vectors were zeroed by k_band_finish
        if (won) for (int p = 0; p < PT.R; ++p) PT.exch[p][(size_t)gain_off + i] = got;
        return;
    }
    gain[i] = got;
}

// Peer-memory mode, lattice state: k_band_finish_move_claim and k_band_graze in ONE launch. The graze needs every claim of
// the step filed, so the two halves are separated by a grid-wide barrier (monotonic counter in global memory: the host passes
// the value it has after this launch's blocks have all arrived). Only launched when the grid is at most one block per SM
// (n <= 256 * SMs), so every block is resident and the spin cannot deadlock. Saves one kernel boundary per step on the
// critical path of a step (the two agent kernels are what is left there besides the stencil).
__global__ void __launch_bounds__(256) k_band_fmc_graze(BandGeom G, double agent_gamma, int32_t *xy, double *st, int n,
                                                        const double *__restrict__ act, int *claim_new, uint8_t *gz, int do_finish,
                                                        double *gain_prev, int *claim_prev, double *reward, uint8_t *done,
                                                        int64_t *agents_done_at, LatCells C, PeerTable PT, int gain_off,
                                                        unsigned int *bar, unsigned int target) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = i < n;
    int x = 0, y = 0;
    uint8_t g = 0;
    if (active) {
        double s = st[i];
        x = xy[2 * i]; y = xy[2 * i + 1];
        if (do_finish) {
            const double gp = gain_prev[i];
            gain_prev[i] = 0.0;
            if (gz[i]) {
                s = s + gp;
                const int lr = band_owned_row(G, x);
                if (lr >= 0) claim_prev[(size_t)lr * G.N + y] = 0x7fffffff;
            }
            s = dw_clip01(s);
            reward[i] = s;
            done[i] = s < 0.1;
            agents_done_at[i] += (s < 0.1) ? 0 : 1;
        }
        const int a = (int)act[i] - 1;
        s = s - agent_gamma;
        st[i] = s;
        if (s > 0.0) {
            if (a != 8) {
                const int d = (a & 2) ? 1 : -1;
                if (((a + 1) & 2) == 0) { y += d; y = y < 0 ? y + G.N : (y >= G.N ? y - G.N : y); }
                else { x += d; x = x < 0 ? x + G.N : (x >= G.N ? x - G.N : x); }
                xy[2 * i] = x;
                xy[2 * i + 1] = y;
            }
            if (a > 4) {
                g = 1;
                const int lr = band_owned_row(G, x);
                if (lr >= 0) atomicMin(claim_new + (size_t)lr * G.N + y, i);
            }
        }
        gz[i] = g;
    }
    // every claim of the step is filed once all blocks have passed here
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(bar, 1u);
        while ((int)(*(volatile unsigned int *)bar - target) < 0) { }
        __threadfence();
    }
    __syncthreads();
    if (!active || !g) return;
    double got = 0.0;
    bool won = false, clear = true;
    const int lr = band_owned_row(G, x);
    if (lr >= 0) {
        if (__ldcg(claim_new + (size_t)lr * G.N + y) == i) { got = C.food((size_t)lr * G.pitch + G.c0 + y); won = true; }
        else clear = false;
    }
    if (clear)
        band_row_images(G, x, [&](int r) { band_col_images(G, y, [&](int cc) { C.zero((size_t)r * G.pitch + cc); }); });
    if (won) for (int p = 0; p < PT.R; ++p) PT.exch[p][(size_t)gain_off + i] = got;
}

// Phase 4 (replicated, after the gains were summed over ranks): state += gain, clip, reward/done, lifespan counter;
// also returns the graze claims of this step to "idle".
__global__ void __launch_bounds__(256) k_band_finish(BandGeom G, const int32_t *__restrict__ xy, int *claim, double *st, int n,
                                                     double *gain, int zero_gain, const uint8_t *__restrict__ gz, double *reward,
                                                     uint8_t *done, int64_t *agents_done_at) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = st[i];
    const double g = gain[i];
    if (zero_gain) gain[i] = 0.0;                        // peer-memory mode: only winners write, so the vector is re-armed here
    if (gz[i]) {
        s = s + g;
        const int lr = band_owned_row(G, xy[2 * i]);
        if (lr >= 0) claim[(size_t)lr * G.N + xy[2 * i + 1]] = 0x7fffffff;
    }
    s = dw_clip01(s);
    st[i] = s;
    reward[i] = s;
    done[i] = s < 0.1;
    agents_done_at[i] += (s < 0.1) ? 0 : 1;
}

// ---- ghost maintenance ---------------------------------------------------------------------------------------------
// The stencil kernels write the ghost COLUMNS of the rows they produce, so a ghost ROW is a copy of a whole stored row
// (pitch words): from the band's own edge rows on a single-rank torus, from the neighbours otherwise.
__global__ void __launch_bounds__(256) k_band_ghost_rows_wrap(BandGeom G, uint32_t *lat) {   // single-rank torus: R == N
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= G.pitch) return;
    lat[c] = lat[(size_t)G.R * G.pitch + c];
    lat[(size_t)(G.R + 1) * G.pitch + c] = lat[(size_t)G.pitch + c];
}

// ---- first step: literal forward from the off-lattice fp64 cover planes [(R+2) x N] onto the padded lattice ---------
__global__ void __launch_bounds__(256) k_band_first_step(DevParams P, double SL, BandGeom Gp, const double *__restrict__ l,
                                                         const double *__restrict__ d, BandGeom Gl, uint32_t *__restrict__ lat_out,
                                                         int *stepmax) {
    const size_t total = (size_t)Gp.R * Gp.N;
    uint32_t mx = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / Gp.N), y = (int)(i - (size_t)r * Gp.N);
        const int ys[3] = {y == 0 ? Gp.N - 1 : y - 1, y, y == Gp.N - 1 ? 0 : y + 1};
        double l9[9], d9[9];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const size_t k = (size_t)(r + a) * Gp.pitch + ys[c];     // stored rows r .. r+2 = band rows r-1 .. r+1
                l9[a * 3 + c] = l[k];
                d9[a * 3 + c] = d[k];
            }
        const LitCell o = dw_literal_cell(P, SL, l9, d9);
        const uint32_t q = dw_pack((int)rint(o.nl * 1000.0), (int)rint(o.nd * 1000.0));
        uint32_t *orow = lat_out + (size_t)(r + 1) * Gl.pitch;
        orow[Gl.c0 + y] = q;
        if (y == 0) orow[Gl.c0 + Gl.N] = q;              // ghost columns of the produced row
        if (y == Gl.N - 1) orow[Gl.c0 - 1] = q;
        mx = __vmaxu2(mx, q);
    }
    const unsigned ml = __reduce_max_sync(0xffffffffu, mx & 0xffffu), md = __reduce_max_sync(0xffffffffu, mx >> 16);
    if ((threadIdx.x & 31) == 0) { atomicMax(stepmax, (int)ml); atomicMax(stepmax + 1, (int)md); }
}

// ---- steady state: one stencil step of the band, 64x64 tiles, TMA-staged halo tile ---------------------------------
#define DWT_TILE 64
#define DWT_TILE_ROWS (DWT_TILE + 2)
#define DWT_TILE_PITCH (DWT_TILE + 8)
#define DWT_TILE_BYTES (DWT_TILE_ROWS * DWT_TILE_PITCH * 4)

struct TiledArgs {
    DevParams P;
    FastCoef F;
    StepCoef C;
    uint32_t *out;          // padded lattice written by this step
    int pitch;
    int tiles_x;            // tiles per row
    int tr_first, tr_skip_lo, tr_skip_hi;   // tile rows of this launch: blockIdx -> tr, skipping [tr_skip_lo, tr_skip_hi)
    int N;
    int *stepmax;           // [2] per-species max of the new band (atomicMax)
    unsigned int *slow_count;
};

__device__ __forceinline__ uint32_t dwt_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct RowsTile72 {
    const uint32_t *t;      // first cell of the thread's first row inside the staged tile
    __device__ __forceinline__ Row6 load(int k) const {
        const uint32_t *p = t + k * DWT_TILE_PITCH;
        return dw_make_row(*reinterpret_cast<const uint4 *>(p), p[-1], p[4]);
    }
};
struct StoreGlobal {
    static constexpr bool kMasks = false;
    uint32_t *o;            // first cell of the thread's first output row
    int pitch;
    __device__ __forceinline__ void operator()(int i, uint32_t (&q)[4]) const {
        *reinterpret_cast<uint4 *>(o + (size_t)i * pitch) = make_uint4(q[0], q[1], q[2], q[3]);
    }
};

// rare path, warp-cooperative: for every lane whose 4x4 tile hit the tie filter, lanes 0..15 re-evaluate one cell of that
// tile each from the staged tile; cells on a rounding tie are recomputed in literal order and patched in the output.
// tile_off / out_off: each lane's offsets of its first cell inside the staged tile / the output lattice.
__device__ __noinline__ uint32_t dwt_fix_warp(const TiledArgs *A, const uint32_t *tile, unsigned flagged, uint32_t mx, int tile_off,
                                              long long out_off, int lane) {
    uint32_t extra = 0;
    bool mine = false;
    __syncwarp();
    while (flagged) {
        const int L = __ffs(flagged) - 1;
        flagged &= flagged - 1;
        const int toff = __shfl_sync(0xffffffffu, tile_off, L);
        const long long ooff = __shfl_sync(0xffffffffu, out_off, L);
        if (lane == L) mine = true;
        if (lane < 16) {
            const uint32_t *p = tile + toff + (lane >> 2) * DWT_TILE_PITCH + (lane & 3);
            const uint32_t *q0 = p - DWT_TILE_PITCH, *q2 = p + DWT_TILE_PITCH;
            const uint32_t E = p[-1] + p[1] + q0[0] + q2[0];
            const uint32_t S8 = E + q0[-1] + q0[1] + q2[-1] + q2[1];
            unsigned tiemin = 0xffffffffu;
            uint32_t v = dw_fast_cell(A->F, A->C, p[0], E, S8, &tiemin);
            if (tiemin < A->F.tie_thresh) {
                double l9[9], d9[9];
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const uint32_t w = p[(a - 1) * DWT_TILE_PITCH + (c - 1)];
                        l9[a * 3 + c] = dw_milli(w & 0xffffu);
                        d9[a * 3 + c] = dw_milli(w >> 16);
                    }
                const LitCell lc = dw_literal_cell(A->P, A->C.SL, l9, d9);
                v = dw_pack((int)rint(lc.nl * 1000.0), (int)rint(lc.nd * 1000.0));
                A->out[ooff + (long long)(lane >> 2) * A->pitch + (lane & 3)] = v;
                if (A->slow_count) atomicAdd(A->slow_count, 1u);
            }
            extra = __vmaxu2(extra, v);
        }
    }
    __syncwarp();                                        // patches (made by helper lanes) before the owners re-read their cells
    return __vmaxu2(mine ? 0u : mx, extra);
}

// A CTA walks the tiles blockIdx.x, blockIdx.x + gridDim.x, ... of the launch with TWO staged tiles: the TMA load of the next
// tile is issued before the current one is computed. The default launch is one tile per CTA (gridDim.x = number of tiles);
// the persistent form (gridDim.x = resident CTAs, DW_TILED_PERSISTENT=1) was measured slower, see dwt_stencil_grid.
__device__ __forceinline__ void dwt_issue_tile(const CUtensorMap *tmap, const TiledArgs &A, int t, uint32_t smem_tile, uint32_t bar_a) {
    const int tc = t % A.tiles_x;
    int tr = A.tr_first + t / A.tiles_x;
    if (tr >= A.tr_skip_lo) tr += A.tr_skip_hi - A.tr_skip_lo;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"((uint32_t)DWT_TILE_BYTES) : "memory");
    // box origin (word, row) of the padded lattice: tile columns 64*tc-4 .. 64*tc+67, band rows 64*tr-1 .. 64*tr+64
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_tile), "l"(tmap), "r"(bar_a), "r"(tc * DWT_TILE), "r"(tr * DWT_TILE)
                 : "memory");
}

template <int STAGES>        // 1: the default one-tile-per-CTA launch (19 KB of shared memory); 2: persistent form with prefetch
__global__ void __launch_bounds__(256, 4) k_tiled_step(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ TiledArgs A,
                                                       int n_tiles) {
    constexpr int kStageWords = ((DWT_TILE_BYTES + 127) / 128) * 32;      // a TMA destination must be 128-byte aligned: 19008 -> 19072 B
    __shared__ __align__(128) uint32_t tile[STAGES][kStageWords];
    __shared__ __align__(8) unsigned long long bar[2];
    __shared__ int smax[2];
    const int tid = threadIdx.x, lane = tid & 31;
    const int tx = tid & 15, r0 = (tid >> 4) * 4;
    const uint32_t bar_a[2] = {dwt_smem_u32(&bar[0]), dwt_smem_u32(&bar[1])};
    const uint32_t tile_a[2] = {dwt_smem_u32(tile[0]), dwt_smem_u32(tile[STAGES - 1])};
    int t = blockIdx.x;
    if (tid == 0) {
        smax[0] = 0; smax[1] = 0;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a[0]));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a[1]));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (t < n_tiles) dwt_issue_tile(&tmap, A, t, tile_a[0], bar_a[0]);
    }
    __syncthreads();                                     // barriers initialised (and the first one armed) before anyone polls
    const int tile_off = (1 + r0) * DWT_TILE_PITCH + 4 + 4 * tx;
    uint32_t mx_all = 0;
    for (int k = 0; t < n_tiles; ++k, t += gridDim.x) {
        const int s = STAGES == 2 ? (k & 1) : 0;
        // the other stage was last read in iteration k-1, which ended with a CTA barrier: free to be refilled
        if (STAGES == 2 && tid == 0 && t + (int)gridDim.x < n_tiles) dwt_issue_tile(&tmap, A, t + gridDim.x, tile_a[s ^ 1], bar_a[s ^ 1]);
        {
            const uint32_t parity = (uint32_t)(STAGES == 2 ? (k >> 1) : k) & 1u;
            uint32_t ok = 0;
            while (!ok)
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                             : "=r"(ok) : "r"(bar_a[s]), "r"(parity) : "memory");
        }
        const int tc = t % A.tiles_x;
        int tr = A.tr_first + t / A.tiles_x;
        if (tr >= A.tr_skip_lo) tr += A.tr_skip_hi - A.tr_skip_lo;
        const long long out_off = (long long)(tr * DWT_TILE + 1 + r0) * A.pitch + 4 + tc * DWT_TILE + 4 * tx;
        unsigned tiemin = 0xffffffffu;
        uint32_t mx = dw_tile_core(A.F, A.C, RowsTile72{tile[s] + tile_off}, StoreGlobal{A.out + out_off, A.pitch}, &tiemin);
        const unsigned flagged = __ballot_sync(0xffffffffu, tiemin < A.F.tie_thresh);
        if (flagged) mx = dwt_fix_warp(&A, tile[s], flagged, mx, tile_off, out_off, lane);
        // ghost columns of the produced rows: the two threads of a row group that hold column 0 / N-1 copy their (final) cells
        if (tx == 0 && tc == 0) {
            uint32_t *o = A.out + out_off;
#pragma unroll
            for (int i = 0; i < 4; ++i) o[(size_t)i * A.pitch + A.N] = o[(size_t)i * A.pitch];
        } else if (tx == 15 && tc == A.tiles_x - 1) {
            uint32_t *o = A.out + out_off + 3;
#pragma unroll
            for (int i = 0; i < 4; ++i) o[(size_t)i * A.pitch - A.N] = o[(size_t)i * A.pitch];
        }
        mx_all = __vmaxu2(mx_all, mx);
        __syncthreads();                                 // every read of tile[s] is done before the next iteration refills it
        if (STAGES == 1 && tid == 0 && t + (int)gridDim.x < n_tiles) dwt_issue_tile(&tmap, A, t + gridDim.x, tile_a[0], bar_a[0]);
    }
    const unsigned ml = __reduce_max_sync(0xffffffffu, mx_all & 0xffffu), md = __reduce_max_sync(0xffffffffu, mx_all >> 16);
    if (lane == 0) { atomicMax(&smax[0], (int)ml); atomicMax(&smax[1], (int)md); }
    __syncthreads();
    if (tid == 0) {
        if (smax[0] > 0) atomicMax(A.stepmax, smax[0]);
        if (smax[1] > 0) atomicMax(A.stepmax + 1, smax[1]);
    }
}

// ---- look-ahead decisions (peer-memory mode) ---------------------------------------------------------------------------
// The greedy decision of step j+1 reads the four neighbours of the agent in the lattice step j PRODUCES -- 4 cells per
// agent out of N x R. Waiting for the whole stencil to finish puts decide + its barrier on the critical path of every
// step; instead the owner band re-evaluates just those cells itself from the step's INPUT lattice (post-graze `lat_old`,
// same fast path + literal tie fix-up as k_tiled_step, hence the same values) while the interior tiles are still being
// computed on the main stream. Neighbours that sit in a ghost row belong to the neighbouring band: those come from the
// NEW ghost rows, which the halo push + barrier that precede this kernel on the side stream have delivered.
struct LookArgs {
    DevParams P;
    FastCoef F;
    StepCoef C;
};
__device__ __forceinline__ uint32_t dwt_new_cell(const LookArgs &A, const BandGeom &G, const uint32_t *__restrict__ lat_old,
                                                 const uint32_t *__restrict__ lat_new, int r, int yc) {
    if (r == 0 || r == G.R + 1) return lat_new[(size_t)r * G.pitch + G.c0 + yc];
    const uint32_t *p = lat_old + (size_t)r * G.pitch + G.c0 + yc;
    const uint32_t *q0 = p - G.pitch, *q2 = p + G.pitch;
    const uint32_t E = p[-1] + p[1] + q0[0] + q2[0];
    const uint32_t S8 = E + q0[-1] + q0[1] + q2[-1] + q2[1];
    unsigned tiemin = 0xffffffffu;
    uint32_t v = dw_fast_cell(A.F, A.C, p[0], E, S8, &tiemin);
    if (tiemin < A.F.tie_thresh) {
        double l9[9], d9[9];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const uint32_t w = p[(a - 1) * G.pitch + (c - 1)];
                l9[a * 3 + c] = dw_milli(w & 0xffffu);
                d9[a * 3 + c] = dw_milli(w >> 16);
            }
        const LitCell lc = dw_literal_cell(A.P, A.C.SL, l9, d9);
        v = dw_pack((int)rint(lc.nl * 1000.0), (int)rint(lc.nd * 1000.0));
    }
    return v;
}
__global__ void __launch_bounds__(256) k_band_lookahead_decide(BandGeom G, const uint32_t *__restrict__ lat_old,
                                                               const uint32_t *__restrict__ lat_new, const int32_t *__restrict__ xy, int n,
                                                               int policy, PeerTable PT, const __grid_constant__ LookArgs A,
                                                               unsigned int epoch, unsigned int *timed_out, unsigned int *ticket) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const int x = xy[2 * i], y = xy[2 * i + 1];
        const int lr = band_owned_row(G, x);
        if (lr >= 0) {
            const int ym = y == 0 ? G.N - 1 : y - 1, yp = y == G.N - 1 ? 0 : y + 1;
            // candidates in action order 4..7 = (x, y-1), (x-1, y), (x+1, y), (x, y+1), like k_band_decide_one
            const double food[4] = {dw_food(dwt_new_cell(A, G, lat_old, lat_new, lr, ym)), dw_food(dwt_new_cell(A, G, lat_old, lat_new, lr - 1, y)),
                                    dw_food(dwt_new_cell(A, G, lat_old, lat_new, lr + 1, y)), dw_food(dwt_new_cell(A, G, lat_old, lat_new, lr, yp))};
            const double out = (double)(dw_greedy_pick(food, policy == DW_POLICY_GREEDY) + 1);
            for (int p = 0; p < PT.R; ++p) PT.exch[p][2 * (size_t)n + i] = out;     // owner publishes to every rank
        }
    }
    dw_last_block_barrier(PT, epoch, timed_out, ticket);
}

// ---- materialisation / synthetic reset -------------------------------------------------------------------------------
// full reference grid [7, R, N] of the band, materialised by a literal forward from the post-graze state the last step
// started from: the other lattice buffer, or the fp64 planes when that step was the first one after a reset
struct PreLattice {
    const uint32_t *p;
    __device__ __forceinline__ void load9(const BandGeom &G, int r, int y, double (&l9)[9], double (&d9)[9]) const {
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const uint32_t w = p[(size_t)(r + a) * G.pitch + G.c0 + y + c - 1];
                l9[a * 3 + c] = dw_milli(w & 0xffffu);
                d9[a * 3 + c] = dw_milli(w >> 16);
            }
    }
};
struct PrePlanes {
    const double *l, *d;
    __device__ __forceinline__ void load9(const BandGeom &G, int r, int y, double (&l9)[9], double (&d9)[9]) const {
        const int ys[3] = {y == 0 ? G.N - 1 : y - 1, y, y == G.N - 1 ? 0 : y + 1};
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const size_t k = (size_t)(r + a) * G.pitch + ys[c];
                l9[a * 3 + c] = l[k];
                d9[a * 3 + c] = d[k];
            }
    }
};
template <class Pre>
__global__ void __launch_bounds__(256) k_band_materialise(DevParams P, double SL, BandGeom G, Pre pre, double *__restrict__ out) {
    const size_t RN = (size_t)G.R * G.N;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < RN; i += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / G.N), y = (int)(i - (size_t)r * G.N);
        double l9[9], d9[9];
        pre.load9(G, r, y, l9, d9);
        const LitCell o = dw_literal_cell(P, SL, l9, d9);
        out[i] = dw_round3(o.nb);
        out[RN + i] = dw_round3(o.nl);
        out[2 * RN + i] = dw_round3(o.nd);
        out[3 * RN + i] = dw_round3(o.T);
        out[4 * RN + i] = dw_round3(o.Tl);
        out[5 * RN + i] = dw_round3(o.Td);
        out[6 * RN + i] = 0.0;
    }
}
// lattice covers of the band as fp64 planes [2, R, N] (k/1000 exactly as np.round stores them)
__global__ void __launch_bounds__(256) k_band_covers(BandGeom G, const uint32_t *__restrict__ lat, double *__restrict__ out) {
    const size_t RN = (size_t)G.R * G.N;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < RN; i += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / G.N), y = (int)(i - (size_t)r * G.N);
        const uint32_t w = lat[(size_t)(r + 1) * G.pitch + G.c0 + y];
        out[i] = dw_milli(w & 0xffffu);
        out[RN + i] = dw_milli(w >> 16);
    }
}
// Position-weighted checksum of the band's own cells (exact integer arithmetic on the milli-covers, wraps mod 2^64):
// out[0..3] += {sum kl, sum kd, sum kl * w, sum kd * w}, w = (global_row * 131 + column) % 977 + 1. Additive over bands,
// so the sum over the ranks of a banded world equals the checksum of the same world held as one band.
__global__ void __launch_bounds__(256) k_band_checksum(BandGeom G, const uint32_t *__restrict__ lat, unsigned long long *out) {
    const size_t RN = (size_t)G.R * G.N;
    unsigned long long a[4] = {0, 0, 0, 0};
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < RN; i += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / G.N), y = (int)(i - (size_t)r * G.N);
        const uint32_t w = lat[(size_t)(r + 1) * G.pitch + G.c0 + y];
        const unsigned long long wt = (unsigned long long)((((long long)(G.row0 + r) % G.N) * 131 + y) % 977 + 1);
        a[0] += w & 0xffffu; a[1] += w >> 16; a[2] += (w & 0xffffu) * wt; a[3] += (w >> 16) * wt;
    }
    for (int k = 0; k < 4; ++k) {
        unsigned long long v = a[k];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(out + k, v);
    }
}
// agent stamp of ch4 (forward :454-459): the highest agent index on a cell wins. claim holds INT_MAX outside a stamp.
__global__ void __launch_bounds__(256) k_band_stamp_claim(BandGeom G, const int32_t *__restrict__ xy, int n, int *claim) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int lr = band_owned_row(G, xy[2 * i]);
    if (lr >= 0) atomicMin(claim + (size_t)lr * G.N + xy[2 * i + 1], n - 1 - i);
}
__global__ void __launch_bounds__(256) k_band_stamp_write(BandGeom G, const int32_t *__restrict__ xy, const double *__restrict__ st, int n,
                                                          int *claim, double *ch4, int pass) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int lr = band_owned_row(G, xy[2 * i]);
    if (lr < 0) return;
    const size_t c = (size_t)lr * G.N + xy[2 * i + 1];
    if (pass == 0) { if (claim[c] == n - 1 - i) ch4[(size_t)(lr - 1) * G.N + xy[2 * i + 1]] = st[i]; }
    else claim[c] = 0x7fffffff;
}

// device-side synthetic reset of the band (+ its ghost rows) and of the replicated agents: same distribution as
// initialize_grid / initialize_agents (daisy_world_rl.py:285-302, 173-179), counter RNG keyed by the GLOBAL cell index,
// so any banding of the same world draws the same state.
__global__ void __launch_bounds__(256) k_band_init_random(BandGeom Gp, uint64_t seed, double prop_l, double prop_d, double init_l,
                                                          double init_d, double *l, double *d, int32_t *agent_xy, double *agent_state, int n) {
    const size_t total = (size_t)(Gp.R + 2) * Gp.N;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / Gp.N), y = (int)(i - (size_t)r * Gp.N);
        int x = Gp.row0 + r - 1;
        x = x < 0 ? x + Gp.N : (x >= Gp.N ? x - Gp.N : x);
        const uint64_t gidx = (uint64_t)x * Gp.N + y;
        d[i] = dw_u01(seed, gidx, 0) < prop_d ? init_d * dw_u01(seed, gidx, 1) : 0.0;
        l[i] = dw_u01(seed, gidx, 2) < prop_l ? init_l * dw_u01(seed, gidx, 3) : 0.0;
    }
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < (size_t)n; i += (size_t)gridDim.x * blockDim.x) {
        agent_xy[2 * i] = (int)(dw_u01(seed ^ 0xA5A5A5A5ull, i, 0) * Gp.N);
        agent_xy[2 * i + 1] = (int)(dw_u01(seed ^ 0xA5A5A5A5ull, i, 1) * Gp.N);
        agent_state[i] = 1.0;
    }
}
