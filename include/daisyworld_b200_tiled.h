/*
 * daisyworld_b200_tiled.h -- C ABI of the single-giant-grid path (BASELINE config 5): ONE toroidal N x N RLDaisyWorld
 * (batch_size == 1, daisy/daisy_world_rl.py:13-501) too large for shared memory, optionally split into row bands, one
 * band per GPU.  Same conventions as daisyworld_b200.h (return codes, ownership, stream semantics, no CPU fallback).
 *
 * A handle owns rows [row0, row0 + rows) of the world (rows == N: the whole torus on one GPU) plus one ghost row above
 * and below.  Agents are replicated: every rank holds all n agents (global coordinates) and runs the same agent
 * kernels; only the rank that owns a cell eats from it.
 *
 * One env step (daisy_world_rl.py:475-497) is the phase sequence below.  With several ranks the caller sums two small
 * device vectors across ranks between phases and copies two halo rows to the neighbours after the stencil
 * (therldaisyworld_b200/banded.py does this with torch.distributed: NCCL on GPUs, gloo in the CPU tests):
 *
 *   dwt_decide        Greedy.__call__ (agents/greedy.py:16-30) or replay/none/random, owner rank only -> act[n]
 *                     [ranks > 1 and greedy/antigreedy: all-reduce SUM of act]
 *   dwt_move_graze    update_agents (:181-216): pay agent_gamma, move, graze in agent-index order -> gain[n]
 *                     [ranks > 1: all-reduce SUM of gain]
 *   dwt_finish_agents state += gain, clip (:244), reward/done (:486-492), agents_done_at += !done
 *   dwt_stencil       forward (:434-452) on the band -> new covers (ghost columns included); update_L (:463-473);
 *                     per-step species max. Parts 1 + 2 split it into edge tile rows and interior so that the
 *                     halo exchange can overlap the interior.
 *   ghost rows        toroidal wrap (dwt_halo_wrap, one rank) or the neighbours' edge rows (caller copies whole
 *                     stored rows of N + 8 words).
 * gain[n] and act[n] are adjacent in memory (gain first): finishing step j can be deferred until the decisions of
 * step j + 1 exist, so that ONE all-reduce per step carries both (therldaisyworld_b200/banded.py does that).
 *
 * Constraints of the tiled kernels: N % 64 == 0, rows % 64 == 0, N >= 64, D4-symmetric kernels (the defaults).
 */
#ifndef DAISYWORLD_B200_TILED_H
#define DAISYWORLD_B200_TILED_H

#include "daisyworld_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dwt_handle dwt_handle;

/* Device pointers a multi-rank driver needs (valid until the next dwt_stencil for the halo rows, for the life of the
   handle for the rest). Row pointers address N + 8 packed u32 words (light milli-cover | dark << 16; word 3 and word N + 4
   are the ghost columns). */
typedef struct dwt_ptrs {
    double *act;            /* [n]   action + 1 per agent (0 = not decided here); act == gain + n */
    double *gain;           /* [n]   food eaten per agent on this rank */
    int32_t *stepmax;       /* [4096, 2] per-step max milli-cover of (light, dark) over this band */
    uint32_t *send_top;     /* band row 0 (whole stored row, N + 8 words) -> the rank above's bottom ghost */
    uint32_t *send_bottom;  /* band row rows-1                            -> the rank below's top ghost */
    uint32_t *recv_top;     /* ghost row above the band */
    uint32_t *recv_bottom;  /* ghost row below the band */
} dwt_ptrs;

const char *dwt_last_error(const dwt_handle *h);
/* cfg->dim = N, cfg->n_agents = n, cfg->batch must be 1; rows/row0: the band of this handle; n_ranks: number of bands */
int dwt_create(const dw_config *cfg, int32_t rows, int32_t row0, int32_t n_ranks, dwt_handle **out);
int dwt_destroy(dwt_handle *h);
int dwt_set_config(dwt_handle *h, const dw_config *cfg);
int dwt_set_clock(dwt_handle *h, const dw_clock *clk);
int dwt_get_clock(dwt_handle *h, dw_clock *clk);
int dwt_set_stream(dwt_handle *h, void *cuda_stream);
int dwt_set_epsilon(dwt_handle *h, double epsilon);       /* for DW_POLICY_EPS_GREEDY, see dw_set_epsilon */
int dwt_synchronize(dwt_handle *h);

/* reset(): cover planes of the band INCLUDING its two ghost rows, host fp64 [(rows + 2), N] each (row 0 = world row
   row0 - 1 with toroidal wrap); the state is off the 0.001 lattice until the first step (daisy_world_rl.py:299-302). */
int dwt_upload_covers(dwt_handle *h, const double *light, const double *dark);
/* all n agents, global coordinates: agent_indices[n,2] int64, agent_states[n] */
int dwt_upload_agents(dwt_handle *h, const int64_t *agent_indices, const double *agent_states);
/* device-side synthetic reset (distribution of initialize_grid/initialize_agents, counter RNG keyed by the global cell /
   agent index: every banding of the same world gets the same state) */
int dwt_init_random(dwt_handle *h, uint64_t seed, double light_proportion, double dark_proportion, double initial_al,
                    double initial_ad);

/* phases of one step (see the header comment). actions: host int8 [n] for DW_POLICY_REPLAY, else NULL. */
int dwt_decide(dwt_handle *h, int32_t policy, const int8_t *actions, uint64_t seed);
int dwt_move_graze(dwt_handle *h);
int dwt_finish_agents(dwt_handle *h);
int dwt_stencil(dwt_handle *h, int32_t part);   /* 0: whole band; 1: edge tile rows, then 2: interior tile rows */
int dwt_halo_wrap(dwt_handle *h);      /* one rank: ghost rows from the band's own edge rows */
int dwt_get_ptrs(dwt_handle *h, dwt_ptrs *out);

/* K whole steps on one rank (rows == N): the phases above back to back on the handle's stream, no host sync. */
int dwt_run(dwt_handle *h, int64_t K, int32_t policy, const int8_t *actions /*[K,n] or NULL*/, uint64_t seed);
/* Close a chunk of `K` steps whose per-step maxima sit in stepmax (all-reduced MAX over ranks by the caller when there
   are several): done_at += #steps whose max(light, dark) > 5 (notebook cell 2: grid_done = max <= 0.005);
   first_all_done = first step index of the chunk with the world done, or -1. Clears stepmax. */
int dwt_end_chunk(dwt_handle *h, int32_t K, int32_t *first_all_done);
int dwt_reset_lifespans(dwt_handle *h);
int dwt_get_lifespans(dwt_handle *h, int64_t *done_at /*[1]*/, int64_t *agents_done_at /*[n]*/);

/* ---- peer-memory mode: one process per GPU, no collective on the step path -----------------------------------------
 * Every rank maps every other rank's exchange vector, barrier flags and lattice buffers (CUDA IPC over NVLink/NVSwitch).
 * The owner band of an agent / the winner of a graze stores its result straight into ALL ranks' exchange vectors
 * (exactly one writer per entry: nothing to reduce), a band pushes its edge rows into the neighbours' ghost rows, and
 * ranks meet at device-side flag barriers. dwt_step_p2p is one env step; every rank calls it in lock-step.
 * Setup: dwt_ipc_export on every rank -> all-gather the blobs (any transport) -> dwt_ipc_attach. */
#define DWT_IPC_HANDLE_BYTES 64
#define DWT_PEER_BUFFERS 4          /* lattice buffer 0, lattice buffer 1, exchange vector, barrier flags */
int dwt_ipc_export(dwt_handle *h, void *handles /* [DWT_PEER_BUFFERS][DWT_IPC_HANDLE_BYTES] */);
int dwt_ipc_attach(dwt_handle *h, int32_t rank, int32_t n_ranks, const void *all_handles /* [n_ranks][4][64] */);
/* same-process variant (several bands of one process, tests): raw device pointers [n_ranks][DWT_PEER_BUFFERS] */
int dwt_get_peer_buffers(dwt_handle *h, void **out /* [DWT_PEER_BUFFERS] */);
int dwt_attach_peers(dwt_handle *h, int32_t rank, int32_t n_ranks, void *const *table);
int dwt_step_p2p(dwt_handle *h, int32_t policy, const int8_t *actions /* [n] or NULL */, uint64_t seed);
int dwt_flush_p2p(dwt_handle *h);                       /* finish the agents of the last step */
int dwt_peer_status(dwt_handle *h, int32_t *timed_out); /* 1 if a barrier gave up waiting for a peer */

/* getters (synchronise) */
int dwt_get_agents(dwt_handle *h, int64_t *agent_indices, double *agent_states);
int dwt_get_reward_done(dwt_handle *h, double *reward, uint8_t *done);          /* [n] */
int dwt_get_covers(dwt_handle *h, double *light, double *dark);                 /* band rows, [rows, N] each */
/* env.grid of the band: [7, rows, N] (channels as in the reference; ch4 carries the agent stamp), materialised by a
   literal forward from the post-graze state the last step started from. Needs at least one step. */
int dwt_get_grid(dwt_handle *h, double *grid);
/* Position-weighted checksum of the band's own cells, computed on the device with exact integer arithmetic (mod 2^64):
   out[4] = {sum kl, sum kd, sum kl*w, sum kd*w} over the milli-covers k = 1000 * cover, w = (global_row * 131 + column)
   % 977 + 1. Additive over bands: the sum over the ranks of a banded world equals the checksum of the same world on one
   GPU (bench.py prints it so that the 1/2/4/8-GPU runs can be compared; the reference quantity is env.grid[0, 1:3]). */
int dwt_cover_checksum(dwt_handle *h, uint64_t *out /*[4]*/);
int dwt_debug_slow_count(dwt_handle *h, uint64_t *count);
/* Measurement hook: average duration (microseconds, CUDA events) of `reps` back-to-back launches of the band's stencil
   kernel on the current state -- the compute floor of one step, against which bench.py states the per-step exchange /
   agent overhead. Overwrites the recorded pre-state (dwt_get_grid needs a new step afterwards). */
int dwt_debug_time_stencil(dwt_handle *h, int32_t reps, double *us_per_launch);

#ifdef __cplusplus
}
#endif
#endif /* DAISYWORLD_B200_TILED_H */
