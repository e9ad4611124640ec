// C-ABI of the single-giant-grid path (include/daisyworld_b200_tiled.h). Included at the end of dw_api.cu.
#include "../../include/daisyworld_b200_tiled.h"
#include "dw_tiled.cuh"

static thread_local std::string g_dwt_create_error;

struct dwt_handle {
    dw_config cfg{};
    dw_clock clk{};
    cudaStream_t stream = nullptr;
    std::string err;
    int N = 0, R = 0, row0 = 0, n = 0, n_ranks = 1, pitch = 0;
    // packed padded lattice, ping-pong; cur = buffer holding the current state
    uint32_t *lat[2] = {nullptr, nullptr};
    CUtensorMap tmap[2];
    int cur = 0;
    unsigned int *gridbar = nullptr;     // grid barrier counter of k_band_fmc_graze (monotonic)
    unsigned int gridbar_count = 0;
    int sm_count = 0;
    bool on_lattice = false;     // false: the state lives in the fp64 planes (right after reset)
    bool have_pre = false;       // a step has run: lat[1-cur] (or the planes) hold its post-graze pre-state
    bool pre_is_planes = false;
    double L_last = 0.0;
    double *pl = nullptr, *pd = nullptr;       // fp64 cover planes [(R+2) x N] (off-lattice reset state)
    int *claim = nullptr;                      // [(R+2) x N] graze claims, INT_MAX when idle
    int32_t *agent_xy = nullptr;
    // exch = [gain1 | gain0 | act], n doubles each. gain = gain0 and act are adjacent (one all-reduce covers both in NCCL
    // mode); peer-memory mode alternates between gain0 and gain1 by step parity.
    double *agent_state = nullptr, *exch = nullptr, *reward = nullptr;
    double *act = nullptr, *gain = nullptr;
    // peer-memory mode
    PeerTable pt{};
    unsigned int *flags = nullptr, *timed_out = nullptr, *ticket = nullptr;
    int *claim2 = nullptr;                     // second claim array (claims alternate by step parity in peer-memory mode)
    unsigned int epoch = 0;
    uint32_t *peer_lat[2][DWT_MAX_RANKS] = {};
    std::vector<void *> ipc_opened;
    int gain_parity = 0, pending_parity = 0;
    unsigned int decide_epoch = 0;             // != 0: the next dwt_decide launch ends in a barrier with this epoch
    bool p2p_gain_pending = false;
    // peer-memory mode, two-stream step: side stream for the halo push / barriers / look-ahead decisions, which run under
    // the interior tiles of the stencil on the main stream
    cudaStream_t stream_b = nullptr;
    cudaEvent_t ev_edge = nullptr, ev_side = nullptr;
    bool look_valid = false;                   // act[] already holds the decisions of the NEXT step (k_band_lookahead_decide)
    int look_policy = -1;
    int64_t look_step = -1;
    int overlap = 1;                           // DW_P2P_OVERLAP=0: the one-stream sequence of round 1 (for comparison)
    cudaStream_t stencil_stream = nullptr;     // != nullptr: dwt_stencil launches there instead of h->stream (edge tiles of the two-stream step)
    bool step_open = false;      // dwt_stencil(part 1) done, part 2 pending
    uint8_t *gz = nullptr, *done = nullptr;
    int8_t *replay = nullptr;
    size_t replay_cap = 0;
    int64_t *agents_done_at = nullptr;
    int64_t done_at = 0;
    double epsilon = 0.0;                      // Greedy.epsilon of DW_POLICY_EPS_GREEDY
    int *stepmax = nullptr;                    // [DW_FUSED_MAX_STEPS, 2]
    int chunk_j = 0;                           // steps recorded in stepmax since the last dwt_end_chunk
    unsigned int *slow_count = nullptr;
    double *scratch = nullptr;                 // [7 x R x N] materialisation buffer (lazy)
    unsigned long long *csum = nullptr;        // [4] dwt_cover_checksum accumulator (lazy)
};

static int dwt_fail(dwt_handle *h, int code, const char *what, const char *detail) {
    std::string m = std::string(what) + ": " + detail;
    if (h) h->err = m; else g_dwt_create_error = m;
    return code;
}
#define DWT_TRY(h, expr)                                                                                  \
    do {                                                                                                  \
        cudaError_t e__ = (expr);                                                                         \
        if (e__ != cudaSuccess) return dwt_fail((h), DW_E_CUDA, #expr, cudaGetErrorString(e__));          \
    } while (0)
#define DWT_LAUNCHED(h) DWT_TRY((h), cudaGetLastError())

extern "C" const char *dwt_last_error(const dwt_handle *h) { return h ? h->err.c_str() : g_dwt_create_error.c_str(); }

static BandGeom dwt_geom_lat(const dwt_handle *h) { return BandGeom{h->N, h->R, h->row0, h->pitch, 4}; }
static BandGeom dwt_geom_planes(const dwt_handle *h) { return BandGeom{h->N, h->R, h->row0, h->N, 0}; }

static DevParams dwt_params(const dwt_handle *h) { return make_params_cfg(h->cfg); }

typedef CUresult (*dwt_encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int dwt_make_tmaps(dwt_handle *h) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    DWT_TRY(h, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn || q != cudaDriverEntryPointSuccess) return dwt_fail(h, DW_E_CUDA, "cuTensorMapEncodeTiled", "driver entry point not found");
    for (int i = 0; i < 2; ++i) {
        const cuuint64_t dims[2] = {(cuuint64_t)h->pitch, (cuuint64_t)(h->R + 2)};
        const cuuint64_t strides[1] = {(cuuint64_t)h->pitch * sizeof(uint32_t)};
        const cuuint32_t box[2] = {DWT_TILE_PITCH, DWT_TILE_ROWS};
        const cuuint32_t estr[2] = {1, 1};
        CUresult r = ((dwt_encode_fn)fn)(&h->tmap[i], CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, h->lat[i], dims, strides, box, estr,
                                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return dwt_fail(h, DW_E_CUDA, "cuTensorMapEncodeTiled", "encode failed");
    }
    return DW_OK;
}

extern "C" int dwt_create(const dw_config *cfg, int32_t rows, int32_t row0, int32_t n_ranks, dwt_handle **out) {
    if (!out) return dwt_fail(nullptr, DW_E_INVALID, "dwt_create", "out is NULL");
    *out = nullptr;
    if (!cfg || cfg->batch != 1 || cfg->n_agents < 0) return dwt_fail(nullptr, DW_E_INVALID, "dwt_create", "batch must be 1, n_agents >= 0");
    const int N = cfg->dim;
    if (N < 64 || N % 64 || rows < 64 || rows % 64 || rows > N || row0 < 0 || row0 >= N || n_ranks < 1)
        return dwt_fail(nullptr, DW_E_UNSUPPORTED, "dwt_create", "tiled path needs N % 64 == 0 and rows % 64 == 0 (64 <= rows <= N)");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return dwt_fail(nullptr, DW_E_CUDA, "dwt_create", e != cudaSuccess ? cudaGetErrorString(e) : "no CUDA device (there is no CPU fallback)");
    if (cfg->device < 0 || cfg->device >= ndev) return dwt_fail(nullptr, DW_E_INVALID, "dwt_create", "bad device ordinal");
    e = cudaSetDevice(cfg->device);
    if (e != cudaSuccess) return dwt_fail(nullptr, DW_E_CUDA, "cudaSetDevice", cudaGetErrorString(e));
    if (!dw_fast_path_cfg_ok(*cfg))
        return dwt_fail(nullptr, DW_E_UNSUPPORTED, "dwt_create", "tiled path needs D4-symmetric kernels, a zero-centre uniform albedo kernel and g > 0");
    dwt_handle *h = new dwt_handle();
    h->cfg = *cfg;
    h->N = N; h->R = rows; h->row0 = row0; h->n = cfg->n_agents; h->n_ranks = n_ranks; h->pitch = N + 8;
    const size_t n = h->n ? h->n : 1, padded = (size_t)(rows + 2) * h->pitch, planes = (size_t)(rows + 2) * N;
    auto alloc = [&](void **p, size_t bytes) { return cudaMalloc(p, bytes) == cudaSuccess && cudaMemset(*p, 0, bytes) == cudaSuccess; };
    bool ok = alloc((void **)&h->lat[0], padded * 4) && alloc((void **)&h->lat[1], padded * 4) && alloc((void **)&h->claim, planes * 4) &&
              alloc((void **)&h->agent_xy, n * 8) && alloc((void **)&h->agent_state, n * 8) && alloc((void **)&h->exch, 3 * n * 8) &&
              alloc((void **)&h->flags, (DWT_MAX_RANKS + 2) * 4) &&
              alloc((void **)&h->reward, n * 8) && alloc((void **)&h->gz, n) && alloc((void **)&h->done, n) &&
              alloc((void **)&h->agents_done_at, n * 8) && alloc((void **)&h->stepmax, (size_t)DW_FUSED_MAX_STEPS * 2 * 4) &&
              alloc((void **)&h->slow_count, 4) && alloc((void **)&h->replay, n) &&
              alloc((void **)&h->gridbar, 4);        // here, not lazily: a cudaMalloc between steps synchronises the whole device
    if (ok) h->replay_cap = n;
    if (ok) ok = cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, cfg->device) == cudaSuccess;
    if (ok) {
        // With lazy module loading the FIRST launch of a kernel loads its code, which can synchronise the whole device. The
        // merged agent kernel is first launched at step 2 of a peer-memory run, when a peer band may already be spinning at a
        // flag barrier: with several bands of one process on ONE device (the test harness) the two would wait for each other
        // until the barrier's timeout. Load it here.
        cudaFuncAttributes fa;
        ok = cudaFuncGetAttributes(&fa, k_band_fmc_graze) == cudaSuccess;
    }
    if (ok) ok = cudaMemset(h->claim, 0x7f, planes * 4) == cudaSuccess;       // 0x7f7f7f7f: "idle" (> any agent index)
    h->gain = h->exch ? h->exch + n : nullptr;
    h->act = h->exch ? h->exch + 2 * n : nullptr;
    h->timed_out = h->flags ? h->flags + DWT_MAX_RANKS : nullptr;
    h->ticket = h->flags ? h->flags + DWT_MAX_RANKS + 1 : nullptr;
    if (!ok) {
        g_dwt_create_error = std::string("dwt_create: device allocation failed: ") + cudaGetErrorString(cudaGetLastError());
        dwt_destroy(h);
        return DW_E_CUDA;
    }
    int rc = dwt_make_tmaps(h);
    if (rc) { g_dwt_create_error = h->err; dwt_destroy(h); return rc; }
    h->clk.L = 0.75; h->clk.min_L = 0.75; h->clk.max_L = 1.5; h->clk.ramp_period = 512;
    h->clk.dL = (h->clk.max_L - h->clk.min_L) / 512.0;
    *out = h;
    return DW_OK;
}

extern "C" int dwt_destroy(dwt_handle *h) {
    if (!h) return DW_OK;
    cudaSetDevice(h->cfg.device);
    cudaStreamSynchronize(h->stream);
    for (void *p : h->ipc_opened) cudaIpcCloseMemHandle(p);
    if (h->stream_b) { cudaStreamSynchronize(h->stream_b); cudaStreamDestroy(h->stream_b); }
    if (h->ev_edge) cudaEventDestroy(h->ev_edge);
    if (h->ev_side) cudaEventDestroy(h->ev_side);
    void *ptrs[] = {h->lat[0], h->lat[1], h->pl, h->pd, h->claim, h->claim2, h->agent_xy, h->agent_state, h->exch, h->flags, h->reward, h->gz,
                    h->done, h->replay, h->agents_done_at, h->stepmax, h->slow_count, h->scratch, h->csum, h->gridbar};
    for (void *p : ptrs) if (p) cudaFree(p);
    delete h;
    return DW_OK;
}

extern "C" int dwt_set_config(dwt_handle *h, const dw_config *cfg) {
    if (!h || !cfg) return DW_E_INVALID;
    if (cfg->dim != h->N || cfg->n_agents != h->n || cfg->device != h->cfg.device || cfg->batch != 1)
        return dwt_fail(h, DW_E_INVALID, "dwt_set_config", "shapes/device of a handle are fixed");
    h->cfg = *cfg;
    return DW_OK;
}
extern "C" int dwt_set_clock(dwt_handle *h, const dw_clock *clk) { if (!h || !clk) return DW_E_INVALID; h->clk = *clk; return DW_OK; }
extern "C" int dwt_get_clock(dwt_handle *h, dw_clock *clk) { if (!h || !clk) return DW_E_INVALID; *clk = h->clk; return DW_OK; }
extern "C" int dwt_set_epsilon(dwt_handle *h, double epsilon) {
    if (!h || !(epsilon >= 0.0 && epsilon <= 1.0)) return DW_E_INVALID;
    h->epsilon = epsilon;
    return DW_OK;
}
extern "C" int dwt_set_stream(dwt_handle *h, void *s) { if (!h) return DW_E_INVALID; h->stream = (cudaStream_t)s; return DW_OK; }
extern "C" int dwt_synchronize(dwt_handle *h) {
    if (!h) return DW_E_INVALID;
    DWT_TRY(h, cudaSetDevice(h->cfg.device));
    DWT_TRY(h, cudaStreamSynchronize(h->stream));
    return DW_OK;
}

static int dwt_ensure_planes(dwt_handle *h) {
    const size_t bytes = (size_t)(h->R + 2) * h->N * sizeof(double);
    if (!h->pl) DWT_TRY(h, cudaMalloc((void **)&h->pl, bytes));
    if (!h->pd) DWT_TRY(h, cudaMalloc((void **)&h->pd, bytes));
    return DW_OK;
}
static void dwt_state_reset(dwt_handle *h) {
    h->on_lattice = false;
    h->have_pre = false;
    h->chunk_j = 0;
    h->look_valid = false;
}

extern "C" int dwt_upload_covers(dwt_handle *h, const double *light, const double *dark) {
    if (!h || !light || !dark) return DW_E_INVALID;
    DWT_TRY(h, cudaSetDevice(h->cfg.device));
    int rc = dwt_ensure_planes(h);
    if (rc) return rc;
    const size_t bytes = (size_t)(h->R + 2) * h->N * sizeof(double);
    DWT_TRY(h, cudaMemcpyAsync(h->pl, light, bytes, cudaMemcpyHostToDevice, h->stream));
    DWT_TRY(h, cudaMemcpyAsync(h->pd, dark, bytes, cudaMemcpyHostToDevice, h->stream));
    DWT_TRY(h, cudaMemsetAsync(h->stepmax, 0, (size_t)DW_FUSED_MAX_STEPS * 2 * 4, h->stream));
    dwt_state_reset(h);
    return DW_OK;
}

extern "C" int dwt_upload_agents(dwt_handle *h, const int64_t *agent_indices, const double *agent_states) {
    if (!h) return DW_E_INVALID;
    DWT_TRY(h, cudaSetDevice(h->cfg.device));
    const size_t n = h->n;
    if (!n) return DW_OK;
    if (agent_indices) {
        std::vector<int32_t> xy(n * 2);
        for (size_t i = 0; i < xy.size(); ++i) {
            int64_t v = agent_indices[i] % h->N;
            xy[i] = (int32_t)(v < 0 ? v + h->N : v);
        }
        DWT_TRY(h, cudaMemcpyAsync(h->agent_xy, xy.data(), xy.size() * 4, cudaMemcpyHostToDevice, h->stream));
        DWT_TRY(h, cudaStreamSynchronize(h->stream));
    }
    if (agent_states) DWT_TRY(h, cudaMemcpyAsync(h->agent_state, agent_states, n * 8, cudaMemcpyHostToDevice, h->stream));
    return DW_OK;
}

extern "C" int dwt_init_random(dwt_handle *h, uint64_t seed, double light_proportion, double dark_proportion, double initial_al,
                               double initial_ad) {
    if (!h) return DW_E_INVALID;
    DWT_TRY(h, cudaSetDevice(h->cfg.device));
    int rc = dwt_ensure_planes(h);
    if (rc) return rc;
    k_band_init_random<<<grid_for((size_t)(h->R + 2) * h->N), 256, 0, h->stream>>>(dwt_geom_planes(h), seed, light_proportion, dark_proportion,
                                                                                  initial_al, initial_ad, h->pl, h->pd, h->agent_xy,
                                                                                  h->agent_state, h->n);
    DWT_LAUNCHED(h);
    DWT_TRY(h, cudaMemsetAsync(h->stepmax, 0, (size_t)DW_FUSED_MAX_STEPS * 2 * 4, h->stream));
    dwt_state_reset(h);
    return DW_OK;
}

static inline int dwt_blocks(int n) { return (n + 255) / 256; }

extern "C" int dwt_decide(dwt_handle *h, int32_t policy, const int8_t *actions, uint64_t seed) {
    if (!h || policy < 0 || policy > DW_POLICY_EPS_GREEDY) return DW_E_INVALID;
    DWT_TRY(h, cudaSetDevice(h->cfg.device));
    if (!h->n) return DW_OK;
    policy = dw_resolve_policy(policy, h->epsilon, seed, (uint32_t)h->clk.step_count);
    if (policy == DW_POLICY_REPLAY) {
        if (!actions) return dwt_fail(h, DW_E_INVALID, "dwt_decide", "REPLAY needs actions[n]");
        if (h->replay_cap < (size_t)h->n) {
            if (h->replay) cudaFree(h->replay);
            h->replay = nullptr;
            DWT_TRY(h, cudaMalloc((void **)&h->replay, h->n));
            h->replay_cap = h->n;
        }
        DWT_TRY(h, cudaMemcpyAsync(h->replay, actions, h->n, cudaMemcpyHostToDevice, h->stream));
    }
    const uint32_t step = (uint32_t)h->clk.step_count;
    if (h->on_lattice)
        k_band_decide<LatCells><<<dwt_blocks(h->n), 256, 0, h->stream>>>(dwt_geom_lat(h), LatCells{h->lat[h->cur]}, h->agent_xy, h->n, policy,
                                                                          h->replay, seed, step, h->pt, h->act, h->decide_epoch, h->timed_out, h->ticket);
    else
        k_band_decide<PlaneCells><<<dwt_blocks(h->n), 256, 0, h->stream>>>(dwt_geom_planes(h), PlaneCells{h->pl, h->pd}, h->agent_xy, h->n,
                                                                            policy, h->replay, seed, step, h->pt, h->act, h->decide_epoch, h->timed_out, h->ticket);
    DWT_LAUNCHED(h);
    return DW_OK;
}

// claim array of a step: always the first one outside peer-memory mode, alternating by step parity inside it
static int *dwt_claim(const dwt_handle *h, int parity) { return (h->pt.on && parity) ? h->claim2 : h->claim; }

static int dwt_graze(dwt_handle *h);

extern "C" int dwt_move_graze(dwt_handle *h) {
    if (!h) return DW_E_INVALID;
    DWT_TRY(h, cudaSetDevice(h->cfg.device));
    if (!h->n) return DW_OK;
    const int nb = dwt_blocks(h->n);
    const BandGeom G = h->on_lattice ? dwt_geom_lat(h) : dwt_geom_planes(h);
    k_band_move_claim<<<nb, 256, 0, h->stream>>>(G, h->cfg.agent_gamma, h->agent_xy, h->agent_state, h->n, h->act,
                                                 dwt_claim(h, h->gain_parity), h->gz);
    DWT_LAUNCHED(h);
    return dwt_graze(h);
}

static int dwt_graze(dwt_handle *h) {
    const int nb = dwt_blocks(h->n);
    const BandGeom G = h->on_lattice ? dwt_geom_lat(h) : dwt_geom_planes(h);
    if (h->on_lattice)
        k_band_graze<LatCells><<<nb, 256, 0, h->stream>>>(G, LatCells{h->lat[h->cur]}, h->agent_xy, h->n, h->gz, dwt_claim(h, h->gain_parity),
                                                           h->gain, h->pt, h->gain_parity ? 0 : h->n);
    else
        k_band_graze<PlaneCells><<<nb, 256, 0, h->stream>>>(G, PlaneCells{h->pl, h->pd}, h->agent_xy, h->n, h->gz, dwt_claim(h, h->gain_parity),
                                                             h->gain, h->pt, h->gain_parity ? 0 : h->n);
    DWT_LAUNCHED(h);
    return DW_OK;
}

extern "C" int dwt_finish_agents(dwt_handle *h) {
    if (!h) return DW_E_INVALID;
    DWT_TRY(h, cudaSetDevice(h->cfg.device));
    if (!h->n) return DW_OK;
    double *gain = h->pt.on ? (h->pending_parity ? h->exch : h->exch + h->n) : h->gain;
    k_band_finish<<<dwt_blocks(h->n), 256, 0, h->stream>>>(dwt_geom_lat(h), h->agent_xy, dwt_claim(h, h->pending_parity), h->agent_state, h->n,
                                                           gain, h->pt.on, h->gz,
                                                           h->reward, h->done, h->agents_done_at);
    DWT_LAUNCHED(h);
    return DW_OK;
}

// part 0: the whole band. part 1 / part 2: the two edge tile rows first, then the interior, so that the caller can send
// the new edge rows to the neighbours while the interior is still being computed.
// Grid of the stencil kernel: one tile per CTA (default). DW_TILED_PERSISTENT=1 launches only the resident CTAs (4 per SM),
// each walking several tiles with two staged TMA tiles -- measured SLOWER on a B200 (16384^2: 854 vs 807 us per launch, 4096^2:
// 64.0 vs 59.5 us): the hardware block scheduler balances 64x64 tiles better than a static stride, and with four CTAs per SM
// the TMA latency of a fresh CTA is already hidden by the other three.
static int dwt_stencil_grid(const dwt_handle *h, int n_tiles) {
    if (!getenv("DW_TILED_PERSISTENT")) return n_tiles;
    const int resident = 4 * (h->sm_count > 0 ? h->sm_count : 148);
    return n_tiles < resident ? n_tiles : resident;
}

extern "C" int dwt_stencil(dwt_handle *h, int32_t part) {
    if (!h || part < 0 || part > 2) return DW_E_INVALID;
    DWT_TRY(h, cudaSetDevice(h->cfg.device));
    if (h->chunk_j >= DW_FUSED_MAX_STEPS) return dwt_fail(h, DW_E_STATE, "dwt_stencil", "call dwt_end_chunk at least every 4096 steps");
    if ((part == 2) != h->step_open) return dwt_fail(h, DW_E_STATE, "dwt_stencil", "part 2 must follow part 1");
    int *smax = h->stepmax + 2 * h->chunk_j;
    cudaStream_t st = h->stencil_stream ? h->stencil_stream : h->stream;
    const int tiles_y = h->R / DWT_TILE;
    if (!h->on_lattice || (part == 2 && h->pre_is_planes)) {
        if (part != 2) {
            if (!h->pl) return dwt_fail(h, DW_E_STATE, "dwt_stencil", "no state uploaded");
            // the reset state is off the 0.001 lattice: first step in literal arithmetic, planes -> lattice (not split)
            k_band_first_step<<<grid_for((size_t)h->R * h->N), 256, 0, st>>>(dwt_params(h), h->cfg.S * h->clk.L, dwt_geom_planes(h),
                                                                                     h->pl, h->pd, dwt_geom_lat(h), h->lat[h->cur], smax);
            DWT_LAUNCHED(h);
            h->on_lattice = true;
            h->pre_is_planes = true;
        }
    } else {
        TiledArgs A{};
        A.P = dwt_params(h);
        make_fast_coef(h->cfg, A.F);
        make_step_coef(h->cfg, h->clk.L, A.C);
        if (part != 2) h->cur = 1 - h->cur;            // lat[cur] is the buffer being written from now on
        A.out = h->lat[h->cur];
        A.pitch = h->pitch;
        A.N = h->N;
        A.tiles_x = h->N / DWT_TILE;
        A.stepmax = smax;
        A.slow_count = h->slow_count;
        int rows_of_tiles = tiles_y;
        A.tr_first = 0; A.tr_skip_lo = tiles_y; A.tr_skip_hi = tiles_y;      // part 0: all tile rows
        if (part == 1) {                                 // tile rows 0 and tiles_y-1
            rows_of_tiles = tiles_y >= 2 ? 2 : 1;
            A.tr_skip_lo = 1; A.tr_skip_hi = tiles_y >= 2 ? tiles_y - 1 : 1;
        } else if (part == 2) {                          // tile rows 1 .. tiles_y-2
            rows_of_tiles = tiles_y >= 2 ? tiles_y - 2 : 0;
            A.tr_first = 1;
        }
        if (rows_of_tiles > 0) {
            const int n_tiles = A.tiles_x * rows_of_tiles;
            const int sgrid = dwt_stencil_grid(h, n_tiles);
            if (sgrid == n_tiles) k_tiled_step<1><<<sgrid, 256, 0, st>>>(h->tmap[1 - h->cur], A, n_tiles);
            else k_tiled_step<2><<<sgrid, 256, 0, st>>>(h->tmap[1 - h->cur], A, n_tiles);
            DWT_LAUNCHED(h);
        }
        if (part != 2) h->pre_is_planes = false;
    }
    if (part == 1) { h->step_open = true; return DW_OK; }
    h->step_open = false;
    h->have_pre = true;
    h->L_last = h->clk.L;
    h->chunk_j += 1;
    update_L(h->clk);
    return DW_OK;
}

extern "C" int dwt_halo_wrap(dwt_handle *h) {
    if (!h) return DW_E_INVALID;
    DWT_TRY(h, cudaSetDevice(h->cfg.device));
    if (h->R != h->N) return dwt_fail(h, DW_E_STATE, "dwt_halo_wrap", "only for a handle that owns the whole torus");
    if (!h->on_lattice) return DW_OK;
    k_band_ghost_rows_wrap<<<dwt_blocks(h->pitch), 256, 0, h->stream>>>(dwt_geom_lat(h), h->lat[h->cur]);
    DWT_LAUNCHED(h);
    return DW_OK;
}

extern "C" int dwt_get_ptrs(dwt_handle *h, dwt_ptrs *out) {
    if (!h || !out) return DW_E_INVALID;
    uint32_t *L = h->lat[h->cur];
    out->act = h->act;
    out->gain = h->gain;
    out->stepmax = h->stepmax;
    out->send_top = L + (size_t)1 * h->pitch;
    out->send_bottom = L + (size_t)h->R * h->pitch;
    out->recv_top = L;
    out->recv_bottom = L + (size_t)(h->R + 1) * h->pitch;
    return DW_OK;
}

extern "C" int dwt_run(dwt_handle *h, int64_t K, int32_t policy, const int8_t *actions, uint64_t seed) {
    if (!h || K < 0) return DW_E_INVALID;
    if (h->R != h->N) return dwt_fail(h, DW_E_STATE, "dwt_run", "only for a handle that owns the whole torus (use the phase calls for bands)");
    for (int64_t j = 0; j < K; ++j) {
        int rc = dwt_decide(h, policy, actions ? actions + (size_t)j * h->n : nullptr, seed);
        if (!rc) rc = dwt_move_graze(h);
        if (!rc) rc = dwt_finish_agents(h);
        if (!rc) rc = dwt_stencil(h, 0);
        if (!rc) rc = dwt_halo_wrap(h);
        if (rc) return rc;
    }
    return DW_OK;
}

extern "C" int dwt_end_chunk(dwt_handle *h, int32_t K, int32_t *first_all_done) {
    if (!h || K < 0 || K > h->chunk_j) return DW_E_INVALID;
    DWT_TRY(h, cudaSetDevice(h->cfg.device));
    std::vector<int> m((size_t)2 * (K ? K : 1));
    if (K) DWT_TRY(h, cudaMemcpyAsync(m.data(), h->stepmax, (size_t)2 * K * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    DWT_TRY(h, cudaMemsetAsync(h->stepmax, 0, (size_t)2 * h->chunk_j * sizeof(int), h->stream));
    DWT_TRY(h, cudaStreamSynchronize(h->stream));
    int first = -1;
    for (int j = 0; j < K; ++j) {
        const bool grid_done = (m[2 * j] > m[2 * j + 1] ? m[2 * j] : m[2 * j + 1]) <= 5;
        if (!grid_done) h->done_at += 1;
        else if (first < 0) first = j;
    }
    if (first_all_done) *first_all_done = first;
    h->chunk_j = 0;
    return DW_OK;
}

extern "C" int dwt_reset_lifespans(dwt_handle *h) {
    if (!h) return DW_E_INVALID;
    DWT_TRY(h, cudaSetDevice(h->cfg.device));
    h->done_at = 0;
    if (h->n) DWT_TRY(h, cudaMemsetAsync(h->agents_done_at, 0, (size_t)h->n * 8, h->stream));
    return DW_OK;
}

extern "C" int dwt_get_lifespans(dwt_handle *h, int64_t *done_at, int64_t *agents_done_at) {
    if (!h) return DW_E_INVALID;
    DWT_TRY(h, cudaSetDevice(h->cfg.device));
    if (done_at) *done_at = h->done_at;
    if (agents_done_at && h->n) DWT_TRY(h, cudaMemcpyAsync(agents_done_at, h->agents_done_at, (size_t)h->n * 8, cudaMemcpyDeviceToHost, h->stream));
    DWT_TRY(h, cudaStreamSynchronize(h->stream));
    return DW_OK;
}

extern "C" int dwt_get_agents(dwt_handle *h, int64_t *agent_indices, double *agent_states) {
    if (!h) return DW_E_INVALID;
    DWT_TRY(h, cudaSetDevice(h->cfg.device));
    const size_t n = h->n;
    if (!n) return DW_OK;
    std::vector<int32_t> xy(n * 2);
    if (agent_indices) DWT_TRY(h, cudaMemcpyAsync(xy.data(), h->agent_xy, n * 8, cudaMemcpyDeviceToHost, h->stream));
    if (agent_states) DWT_TRY(h, cudaMemcpyAsync(agent_states, h->agent_state, n * 8, cudaMemcpyDeviceToHost, h->stream));
    DWT_TRY(h, cudaStreamSynchronize(h->stream));
    if (agent_indices) for (size_t i = 0; i < xy.size(); ++i) agent_indices[i] = xy[i];
    return DW_OK;
}

extern "C" int dwt_get_reward_done(dwt_handle *h, double *reward, uint8_t *done) {
    if (!h) return DW_E_INVALID;
    DWT_TRY(h, cudaSetDevice(h->cfg.device));
    if (reward && h->n) DWT_TRY(h, cudaMemcpyAsync(reward, h->reward, (size_t)h->n * 8, cudaMemcpyDeviceToHost, h->stream));
    if (done && h->n) DWT_TRY(h, cudaMemcpyAsync(done, h->done, (size_t)h->n, cudaMemcpyDeviceToHost, h->stream));
    DWT_TRY(h, cudaStreamSynchronize(h->stream));
    return DW_OK;
}

static int dwt_ensure_scratch(dwt_handle *h) {
    if (!h->scratch) DWT_TRY(h, cudaMalloc((void **)&h->scratch, (size_t)7 * h->R * h->N * sizeof(double)));
    return DW_OK;
}

extern "C" int dwt_get_covers(dwt_handle *h, double *light, double *dark) {
    if (!h || !light || !dark) return DW_E_INVALID;
    DWT_TRY(h, cudaSetDevice(h->cfg.device));
    const size_t RN = (size_t)h->R * h->N;
    if (!h->on_lattice) {
        if (!h->pl) return dwt_fail(h, DW_E_STATE, "dwt_get_covers", "no state uploaded");
        DWT_TRY(h, cudaMemcpyAsync(light, h->pl + h->N, RN * 8, cudaMemcpyDeviceToHost, h->stream));
        DWT_TRY(h, cudaMemcpyAsync(dark, h->pd + h->N, RN * 8, cudaMemcpyDeviceToHost, h->stream));
    } else {
        int rc = dwt_ensure_scratch(h);
        if (rc) return rc;
        k_band_covers<<<grid_for(RN), 256, 0, h->stream>>>(dwt_geom_lat(h), h->lat[h->cur], h->scratch);
        DWT_LAUNCHED(h);
        DWT_TRY(h, cudaMemcpyAsync(light, h->scratch, RN * 8, cudaMemcpyDeviceToHost, h->stream));
        DWT_TRY(h, cudaMemcpyAsync(dark, h->scratch + RN, RN * 8, cudaMemcpyDeviceToHost, h->stream));
    }
    DWT_TRY(h, cudaStreamSynchronize(h->stream));
    return DW_OK;
}

extern "C" int dwt_get_grid(dwt_handle *h, double *grid) {
    if (!h || !grid) return DW_E_INVALID;
    DWT_TRY(h, cudaSetDevice(h->cfg.device));
    if (!h->have_pre) return dwt_fail(h, DW_E_STATE, "dwt_get_grid", "needs at least one step since the last reset");
    int rc = dwt_ensure_scratch(h);
    if (rc) return rc;
    const size_t RN = (size_t)h->R * h->N;
    const DevParams P = dwt_params(h);
    if (h->pre_is_planes)
        k_band_materialise<PrePlanes><<<grid_for(RN), 256, 0, h->stream>>>(P, h->cfg.S * h->L_last, dwt_geom_planes(h), PrePlanes{h->pl, h->pd}, h->scratch);
    else
        k_band_materialise<PreLattice><<<grid_for(RN), 256, 0, h->stream>>>(P, h->cfg.S * h->L_last, dwt_geom_lat(h), PreLattice{h->lat[1 - h->cur]},
                                                                            h->scratch);
    DWT_LAUNCHED(h);
    if (h->n) {
        const int nb = dwt_blocks(h->n);
        const BandGeom G = dwt_geom_lat(h);
        k_band_stamp_claim<<<nb, 256, 0, h->stream>>>(G, h->agent_xy, h->n, h->claim);
        k_band_stamp_write<<<nb, 256, 0, h->stream>>>(G, h->agent_xy, h->agent_state, h->n, h->claim, h->scratch + 4 * RN, 0);
        k_band_stamp_write<<<nb, 256, 0, h->stream>>>(G, h->agent_xy, h->agent_state, h->n, h->claim, h->scratch + 4 * RN, 1);
        DWT_LAUNCHED(h);
    }
    DWT_TRY(h, cudaMemcpyAsync(grid, h->scratch, 7 * RN * 8, cudaMemcpyDeviceToHost, h->stream));
    DWT_TRY(h, cudaStreamSynchronize(h->stream));
    return DW_OK;
}

// Measurement hook: the stencil kernel of one step over the whole band (k_tiled_step, all tile rows), `reps` launches back to
// back on the current state, timed with CUDA events -> *us_per_launch. The output goes to the buffer that holds the
// pre-state of the last step, so the lazily materialised grid / diagnostics of that step are gone afterwards (have_pre is
// cleared); covers, agents, clock and counters are untouched.
extern "C" int dwt_debug_time_stencil(dwt_handle *h, int32_t reps, double *us_per_launch) {
    if (!h || reps < 1 || !us_per_launch) return DW_E_INVALID;
    DWT_TRY(h, cudaSetDevice(h->cfg.device));
    if (!h->on_lattice) return dwt_fail(h, DW_E_STATE, "dwt_debug_time_stencil", "the state is off the lattice until the first step has run");
    TiledArgs A{};
    A.P = dwt_params(h);
    make_fast_coef(h->cfg, A.F);
    make_step_coef(h->cfg, h->clk.L, A.C);
    A.out = h->lat[1 - h->cur];
    A.pitch = h->pitch;
    A.N = h->N;
    A.tiles_x = h->N / DWT_TILE;
    if (!h->csum) DWT_TRY(h, cudaMalloc((void **)&h->csum, 4 * sizeof(unsigned long long)));
    A.stepmax = reinterpret_cast<int *>(h->csum);           // scratch: the lifespan bookkeeping must not see these launches
    A.slow_count = nullptr;
    const int tiles_y = h->R / DWT_TILE;
    A.tr_first = 0; A.tr_skip_lo = tiles_y; A.tr_skip_hi = tiles_y;
    cudaEvent_t e0, e1;
    DWT_TRY(h, cudaEventCreate(&e0));
    DWT_TRY(h, cudaEventCreate(&e1));
    const int n_tiles = A.tiles_x * tiles_y, sgrid = dwt_stencil_grid(h, n_tiles);
    void (*kern)(const CUtensorMap, const TiledArgs, int) = sgrid == n_tiles ? k_tiled_step<1> : k_tiled_step<2>;
    kern<<<sgrid, 256, 0, h->stream>>>(h->tmap[h->cur], A, n_tiles);      // warm-up
    DWT_TRY(h, cudaEventRecord(e0, h->stream));
    for (int r = 0; r < reps; ++r) kern<<<sgrid, 256, 0, h->stream>>>(h->tmap[h->cur], A, n_tiles);
    DWT_LAUNCHED(h);
    DWT_TRY(h, cudaEventRecord(e1, h->stream));
    DWT_TRY(h, cudaEventSynchronize(e1));
    float ms = 0.f;
    DWT_TRY(h, cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    h->have_pre = false;
    h->look_valid = false;
    *us_per_launch = (double)ms * 1e3 / reps;
    return DW_OK;
}

extern "C" int dwt_cover_checksum(dwt_handle *h, uint64_t *out) {
    if (!h || !out) return DW_E_INVALID;
    DWT_TRY(h, cudaSetDevice(h->cfg.device));
    if (!h->on_lattice) return dwt_fail(h, DW_E_STATE, "dwt_cover_checksum", "the state is off the lattice until the first step has run");
    if (!h->csum) DWT_TRY(h, cudaMalloc((void **)&h->csum, 4 * sizeof(unsigned long long)));
    unsigned long long *acc = h->csum;
    DWT_TRY(h, cudaMemsetAsync(acc, 0, 4 * sizeof(unsigned long long), h->stream));
    k_band_checksum<<<148 * 8, 256, 0, h->stream>>>(dwt_geom_lat(h), h->lat[h->cur], acc);
    DWT_LAUNCHED(h);
    DWT_TRY(h, cudaMemcpyAsync(out, acc, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
    DWT_TRY(h, cudaStreamSynchronize(h->stream));
    return DW_OK;
}

extern "C" int dwt_debug_slow_count(dwt_handle *h, uint64_t *count) {
    if (!h || !count) return DW_E_INVALID;
    DWT_TRY(h, cudaSetDevice(h->cfg.device));
    unsigned int c = 0;
    DWT_TRY(h, cudaMemcpyAsync(&c, h->slow_count, sizeof(c), cudaMemcpyDeviceToHost, h->stream));
    DWT_TRY(h, cudaStreamSynchronize(h->stream));
    *count = c;
    return DW_OK;
}

// ---- peer-memory mode ----------------------------------------------------------------------------------------------------
enum { DWT_PEER_LAT0 = 0, DWT_PEER_LAT1 = 1, DWT_PEER_EXCH = 2, DWT_PEER_FLAGS = 3 };

extern "C" int dwt_get_peer_buffers(dwt_handle *h, void **out) {
    if (!h || !out) return DW_E_INVALID;
    out[DWT_PEER_LAT0] = h->lat[0]; out[DWT_PEER_LAT1] = h->lat[1]; out[DWT_PEER_EXCH] = h->exch; out[DWT_PEER_FLAGS] = h->flags;
    return DW_OK;
}

extern "C" int dwt_ipc_export(dwt_handle *h, void *handles) {
    if (!h || !handles) return DW_E_INVALID;
    DWT_TRY(h, cudaSetDevice(h->cfg.device));
    void *bufs[DWT_PEER_BUFFERS];
    dwt_get_peer_buffers(h, bufs);
    for (int k = 0; k < DWT_PEER_BUFFERS; ++k) {
        cudaIpcMemHandle_t m;
        DWT_TRY(h, cudaIpcGetMemHandle(&m, bufs[k]));
        static_assert(sizeof(m) == DWT_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t size");
        memcpy((unsigned char *)handles + (size_t)k * DWT_IPC_HANDLE_BYTES, &m, sizeof(m));
    }
    return DW_OK;
}

// table: [n_ranks][DWT_PEER_BUFFERS] device pointers valid in THIS process (own entries may be NULL: the handle's own
// buffers are used for its rank)
extern "C" int dwt_attach_peers(dwt_handle *h, int32_t rank, int32_t n_ranks, void *const *table) {
    if (!h || !table || n_ranks < 2 || n_ranks > DWT_MAX_RANKS || rank < 0 || rank >= n_ranks || n_ranks != h->n_ranks)
        return dwt_fail(h, DW_E_INVALID, "dwt_attach_peers", "2 <= n_ranks <= 8 ranks, matching dwt_create");
    DWT_TRY(h, cudaSetDevice(h->cfg.device));
    void *own[DWT_PEER_BUFFERS];
    dwt_get_peer_buffers(h, own);
    PeerTable T{};
    T.rank = rank; T.R = n_ranks; T.on = 1;
    T.timeout_clocks = 0;
    if (const char *ms = getenv("DW_PEER_TIMEOUT_MS")) T.timeout_clocks = (long long)(atof(ms) * 1.9e6);   // ~1.9 GHz SM clock
    for (int r = 0; r < n_ranks; ++r) {
        void *const *row = table + (size_t)r * DWT_PEER_BUFFERS;
        auto pick = [&](int k) { return r == rank ? own[k] : row[k]; };
        if (!pick(DWT_PEER_LAT0) || !pick(DWT_PEER_LAT1) || !pick(DWT_PEER_EXCH) || !pick(DWT_PEER_FLAGS))
            return dwt_fail(h, DW_E_INVALID, "dwt_attach_peers", "missing peer pointer");
        h->peer_lat[0][r] = (uint32_t *)pick(DWT_PEER_LAT0);
        h->peer_lat[1][r] = (uint32_t *)pick(DWT_PEER_LAT1);
        T.exch[r] = (double *)pick(DWT_PEER_EXCH);
        T.flags[r] = (unsigned int *)pick(DWT_PEER_FLAGS);
    }
    // Force-load every kernel of the step now: with lazy module loading the first launch of a kernel can wait for the
    // device to drain, which never happens while a peer band of the same process spins in a barrier waiting for us.
    {
        cudaFuncAttributes fa;
        const void *ks[] = {(const void *)k_band_decide<LatCells>, (const void *)k_band_decide<PlaneCells>, (const void *)k_band_move_claim,
                            (const void *)k_band_graze<LatCells>, (const void *)k_band_graze<PlaneCells>, (const void *)k_band_finish,
                            (const void *)k_band_first_step, (const void *)k_tiled_step<1>, (const void *)k_tiled_step<2>, (const void *)k_band_push_halo,
                            (const void *)k_band_finish_move_claim, (const void *)k_band_fmc_graze, (const void *)k_band_lookahead_decide,
                            (const void *)k_peer_barrier, (const void *)k_band_covers, (const void *)k_band_materialise<PreLattice>,
                            (const void *)k_band_materialise<PrePlanes>, (const void *)k_band_stamp_claim, (const void *)k_band_stamp_write,
                            (const void *)k_band_ghost_rows_wrap, (const void *)k_band_init_random};
        for (const void *k : ks) DWT_TRY(h, cudaFuncGetAttributes(&fa, k));
    }
    if (!h->claim2) {
        const size_t bytes = (size_t)(h->R + 2) * h->N * sizeof(int);
        DWT_TRY(h, cudaMalloc((void **)&h->claim2, bytes));
        DWT_TRY(h, cudaMemsetAsync(h->claim2, 0x7f, bytes, h->stream));
    }
    DWT_TRY(h, cudaMemsetAsync(h->flags, 0, (DWT_MAX_RANKS + 2) * 4, h->stream));
    DWT_TRY(h, cudaMemsetAsync(h->exch, 0, (size_t)3 * (h->n ? h->n : 1) * 8, h->stream));
    DWT_TRY(h, cudaStreamSynchronize(h->stream));
    h->pt = T;
    h->epoch = 0;
    h->gain_parity = 0;
    h->p2p_gain_pending = false;
    h->look_valid = false;
    if (!h->stream_b) {
        // HIGHEST priority: its small kernels must get CTA slots while thousands of interior-tile CTAs of the main stream
        // are still queued (at equal priority the block scheduler drains the earlier grid first and nothing overlaps)
        int prio_lo = 0, prio_hi = 0;
        DWT_TRY(h, cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        if (const char *pr = getenv("DW_P2P_SIDE_PRIO")) { if (!atoi(pr)) prio_hi = prio_lo; }     // experiments: 0 = equal priority
        DWT_TRY(h, cudaStreamCreateWithPriority(&h->stream_b, cudaStreamNonBlocking, prio_hi));
        DWT_TRY(h, cudaEventCreateWithFlags(&h->ev_edge, cudaEventDisableTiming));
        DWT_TRY(h, cudaEventCreateWithFlags(&h->ev_side, cudaEventDisableTiming));
    }
    if (const char *ov = getenv("DW_P2P_OVERLAP")) h->overlap = atoi(ov);
    return DW_OK;
}

extern "C" int dwt_ipc_attach(dwt_handle *h, int32_t rank, int32_t n_ranks, const void *all_handles) {
    if (!h || !all_handles || n_ranks < 2 || n_ranks > DWT_MAX_RANKS || rank < 0 || rank >= n_ranks) return DW_E_INVALID;
    DWT_TRY(h, cudaSetDevice(h->cfg.device));
    std::vector<void *> table((size_t)n_ranks * DWT_PEER_BUFFERS, nullptr);
    for (int r = 0; r < n_ranks; ++r) {
        if (r == rank) continue;
        for (int k = 0; k < DWT_PEER_BUFFERS; ++k) {
            cudaIpcMemHandle_t m;
            memcpy(&m, (const unsigned char *)all_handles + ((size_t)r * DWT_PEER_BUFFERS + k) * DWT_IPC_HANDLE_BYTES, sizeof(m));
            void *p = nullptr;
            DWT_TRY(h, cudaIpcOpenMemHandle(&p, m, cudaIpcMemLazyEnablePeerAccess));
            h->ipc_opened.push_back(p);
            table[(size_t)r * DWT_PEER_BUFFERS + k] = p;
        }
    }
    return dwt_attach_peers(h, rank, n_ranks, table.data());
}

// One env step of a band in peer-memory mode: no collective, no host synchronisation. Every rank must call it the same
// number of times with the same policy. The agents of the step are finished (state += gain, reward/done) at the start of
// the next step or by dwt_flush_p2p.
//
// Two streams (h->overlap, default): the main stream runs   [decide + barrier, only when no look-ahead decisions exist]
//   finish(j-1)+move+claim(j) -> graze(j) -> stencil of the INTERIOR tile rows;
// the side stream (highest priority, so its CTAs are scheduled ahead of the queued interior tiles), started by an event
// after the graze, runs concurrently
//   stencil of the two EDGE tile rows -> push of the new edge rows into the neighbours' ghost rows + closing barrier of
//   step j -> look-ahead decisions of step j+1 (k_band_lookahead_decide) + barrier,
// and the main stream waits for it before the next step's move. What is left on the critical path of a step besides the
// stencil are the two small agent kernels.
extern "C" int dwt_step_p2p(dwt_handle *h, int32_t policy, const int8_t *actions, uint64_t seed) {
    if (!h) return DW_E_INVALID;
    if (!h->pt.on) return dwt_fail(h, DW_E_STATE, "dwt_step_p2p", "attach the peers first (dwt_ipc_attach / dwt_attach_peers)");
    int rc = DW_OK;
    DWT_TRY(h, cudaSetDevice(h->cfg.device));
    const bool overlap = h->overlap != 0;
    const bool was_on_lattice = h->on_lattice;
    if (h->n) {
        // decisions that read the world are published by the owner ranks: the decide kernel ends in a barrier (raised by
        // its last block) so that everyone has every decision before anyone moves
        const int pol = dw_resolve_policy(policy, h->epsilon, seed, (uint32_t)h->clk.step_count);
        const bool world_policy = pol == DW_POLICY_GREEDY || pol == DW_POLICY_ANTIGREEDY;
        const bool have = h->look_valid && h->look_policy == pol && h->look_step == h->clk.step_count;
        h->look_valid = false;
        if (!have) {
            h->decide_epoch = world_policy ? ++h->epoch : 0u;
            rc = dwt_decide(h, policy, actions, seed);
            h->decide_epoch = 0u;
            if (rc) return rc;
        }
        // finish of the previous step (its gains are complete since its closing barrier) + move + claims, one launch
        const BandGeom G = h->on_lattice ? dwt_geom_lat(h) : dwt_geom_planes(h);
        double *gain_prev = h->pending_parity ? h->exch : h->exch + h->n;
        const int nb = dwt_blocks(h->n);
        if (h->on_lattice && nb <= h->sm_count && !getenv("DW_BAND_SPLIT_AGENT_KERNELS")) {
            // finish(j-1) + move + claim + graze in one launch (grid barrier between claim and graze, see k_band_fmc_graze)
            h->gridbar_count += (unsigned int)nb;
            k_band_fmc_graze<<<nb, 256, 0, h->stream>>>(G, h->cfg.agent_gamma, h->agent_xy, h->agent_state, h->n, h->act,
                                                        dwt_claim(h, h->gain_parity), h->gz, h->p2p_gain_pending ? 1 : 0, gain_prev,
                                                        dwt_claim(h, h->pending_parity), h->reward, h->done, h->agents_done_at,
                                                        LatCells{h->lat[h->cur]}, h->pt, h->gain_parity ? 0 : h->n, h->gridbar,
                                                        h->gridbar_count);
            DWT_LAUNCHED(h);
        } else {
            k_band_finish_move_claim<<<nb, 256, 0, h->stream>>>(G, h->cfg.agent_gamma, h->agent_xy, h->agent_state, h->n, h->act,
                                                                dwt_claim(h, h->gain_parity), h->gz, h->p2p_gain_pending ? 1 : 0,
                                                                gain_prev, dwt_claim(h, h->pending_parity), h->reward, h->done,
                                                                h->agents_done_at);
            DWT_LAUNCHED(h);
            rc = dwt_graze(h);
            if (rc) return rc;
        }
        h->p2p_gain_pending = true;
        h->pending_parity = h->gain_parity;
        h->gain_parity ^= 1;
    }
    const int up = (h->pt.rank + h->pt.R - 1) % h->pt.R, down = (h->pt.rank + 1) % h->pt.R;
    if (!overlap) {
        rc = dwt_stencil(h, 0);
        if (rc) return rc;
        // edge rows into the neighbours' ghost rows, then the closing barrier of the step (same launch)
        h->epoch += 1;
        k_band_push_halo<<<dwt_blocks(h->pitch), 256, 0, h->stream>>>(h->lat[h->cur], h->R, h->pitch,
                                                                      h->peer_lat[h->cur][up] + (size_t)(h->R + 1) * h->pitch,
                                                                      h->peer_lat[h->cur][down], h->pt, h->epoch, h->timed_out, h->ticket);
        DWT_LAUNCHED(h);
        return DW_OK;
    }
    // the coefficients of THIS step (the look-ahead re-evaluates cells of the lattice this step produces)
    LookArgs LA{};
    LA.P = dwt_params(h);
    make_fast_coef(h->cfg, LA.F);
    make_step_coef(h->cfg, h->clk.L, LA.C);
    // agents done: the edge tile rows go to the (high-priority) side stream, the interior tile rows stay on the main
    // stream -- disjoint output rows of the same step, both read the post-graze input buffer
    DWT_TRY(h, cudaEventRecord(h->ev_edge, h->stream));
    DWT_TRY(h, cudaStreamWaitEvent(h->stream_b, h->ev_edge, 0));
    h->stencil_stream = h->stream_b;
    rc = dwt_stencil(h, 1);                              // edge tile rows (or the whole literal first step)
    h->stencil_stream = nullptr;
    if (rc) return rc;
    rc = dwt_stencil(h, 2);                              // interior tile rows; advances the clock
    if (rc) return rc;
    h->epoch += 1;
    k_band_push_halo<<<dwt_blocks(h->pitch), 256, 0, h->stream_b>>>(h->lat[h->cur], h->R, h->pitch,
                                                                    h->peer_lat[h->cur][up] + (size_t)(h->R + 1) * h->pitch,
                                                                    h->peer_lat[h->cur][down], h->pt, h->epoch, h->timed_out, h->ticket);
    DWT_LAUNCHED(h);
    if (h->n && was_on_lattice) {
        const int pol_next = dw_resolve_policy(policy, h->epsilon, seed, (uint32_t)h->clk.step_count);
        if (pol_next == DW_POLICY_GREEDY || pol_next == DW_POLICY_ANTIGREEDY) {
            h->epoch += 1;
            k_band_lookahead_decide<<<dwt_blocks(h->n), 256, 0, h->stream_b>>>(dwt_geom_lat(h), h->lat[1 - h->cur], h->lat[h->cur], h->agent_xy,
                                                                               h->n, pol_next, h->pt, LA, h->epoch, h->timed_out, h->ticket);
            DWT_LAUNCHED(h);
            h->look_valid = true;
        } else if (pol_next == DW_POLICY_NONE || pol_next == DW_POLICY_RANDOM) {
            // policies that do not read the world: every rank fills its own act[] (no exchange, no barrier), off the main stream too
            k_band_decide<LatCells><<<dwt_blocks(h->n), 256, 0, h->stream_b>>>(dwt_geom_lat(h), LatCells{h->lat[h->cur]}, h->agent_xy, h->n, pol_next,
                                                                              h->replay, seed, (uint32_t)h->clk.step_count, h->pt, h->act, 0u,
                                                                              h->timed_out, h->ticket);
            DWT_LAUNCHED(h);
            h->look_valid = true;
        }
        if (h->look_valid) {
            h->look_policy = pol_next;
            h->look_step = h->clk.step_count;
        }
    }
    DWT_TRY(h, cudaEventRecord(h->ev_side, h->stream_b));
    DWT_TRY(h, cudaStreamWaitEvent(h->stream, h->ev_side, 0));
    return DW_OK;
}

extern "C" int dwt_flush_p2p(dwt_handle *h) {
    if (!h) return DW_E_INVALID;
    if (h->pt.on && h->p2p_gain_pending) {
        int rc = dwt_finish_agents(h);
        if (rc) return rc;
        h->p2p_gain_pending = false;
    }
    return DW_OK;
}

extern "C" int dwt_peer_status(dwt_handle *h, int32_t *timed_out) {
    if (!h || !timed_out) return DW_E_INVALID;
    DWT_TRY(h, cudaSetDevice(h->cfg.device));
    unsigned int v = 0;
    DWT_TRY(h, cudaMemcpyAsync(&v, h->timed_out, 4, cudaMemcpyDeviceToHost, h->stream));
    DWT_TRY(h, cudaStreamSynchronize(h->stream));
    *timed_out = (int32_t)v;
    return DW_OK;
}
