// dw_run / dw_run_chunk: K steps with an on-device policy and the notebook lifespan counters.
// Included at the end of dw_api.cu.

// ---- fused lattice path: host side -----------------------------------------------------------------------
static size_t fused_smem_bytes(int N, int n) {
    const size_t NN = (size_t)N * N;
    return 2 * NN * sizeof(uint32_t) + (size_t)n * sizeof(double) + 3 * (size_t)n * sizeof(int) + 8 * sizeof(int);
}

// The fast path needs D4-symmetric 3x3 kernels (centre / edge / corner classes) and a uniform zero-centre
// adjacent kernel; anything else (and worlds too large for shared memory) runs through the materialising kernels.
static bool dw_fast_path_cfg_ok(const dw_config &c) {
    const double *w = c.daisy_kernel, *a = c.adjacent_kernel;
    const bool wsym = w[0] == w[2] && w[0] == w[6] && w[0] == w[8] && w[1] == w[3] && w[1] == w[5] && w[1] == w[7];
    bool asym = a[4] == 0.0;
    for (int i = 0; i < 9; ++i) if (i != 4 && a[i] != a[0]) asym = false;
    return wsym && asym && c.g > 0.0;                  // the fast path scales X by g^2 and T by sqrt(g)
}
static bool dw_fused_supported(const dw_handle *h) {
    const dw_config &c = h->cfg;
    if (!dw_fast_path_cfg_ok(c)) return false;
    if (c.n_agents > DW_FUSED_MAX_AGENTS) return false;
    if (getenv("DW_DISABLE_FUSED")) return false;
    return fused_smem_bytes(c.dim, c.n_agents) <= 200 * 1024;
}

static void make_fast_coef(const dw_config &c, FastCoef &F) {
    F.w0 = c.daisy_kernel[4];
    F.w12 = c.daisy_kernel[1] - c.daisy_kernel[0];
    F.w2 = c.daisy_kernel[0];
    F.dtp = c.dt * c.p;
    F.dtm = c.dt / 1000.0;
    F.dtg = c.dt * c.gamma;
    const double cl = (c.albedo_light - c.albedo_bare) / 1000.0, cd = (c.albedo_dark - c.albedo_bare) / 1000.0;
    const double g2 = c.g * c.g;        // X' = g^2 X, T' = sqrt(g) T (see FastCoef)
    F.xk_l = g2 * ((c.q2 - c.q) * cl);
    F.xk_d = g2 * ((c.q2 - c.q) * cd);
    F.xdd = g2 * (c.q2 * (c.albedo_light - c.albedo_dark));
    F.topt = sqrt(c.g) * c.temp_optimal;
    const double Al0 = c.albedo_bare * c.p;
    F.t0 = -g2 * (c.q2 * (Al0 - c.albedo_light));
    F.tk_l = -g2 * (c.q2 * cl);
    F.tk_d = -g2 * (c.q2 * cd);
    // Tie-filter half-width from the error budget: the fast fourth root is good to 3e-12 relative
    // (asserted in the tests; measured maximum 2.2e-12), T stays below 400 K and within max(|Topt-150|, |Topt-400|) of the
    // optimum, so |dx| <= 1000 * dt * 2 g |Topt - T| * 400 * 3e-12 milli-cover (both extremes at once: conservative);
    // the filter is 3x that, never below DW_TIE_EPS_MIN.
    const double dT = fmax(fabs(c.temp_optimal - 150.0), fabs(c.temp_optimal - 400.0));
    const double bound = 1000.0 * fabs(c.dt) * 2.0 * c.g * dT * 400.0 * 3e-12;
    double eps = ceil(3.0 * bound * (double)(1 << DW_FIX_BITS));
    eps = eps < DW_TIE_EPS_MIN ? DW_TIE_EPS_MIN : (eps > DW_TIE_EPS_MAX ? DW_TIE_EPS_MAX : eps);
    F.magic = 6442450944.0 + 0.5 + eps / (double)(1 << DW_FIX_BITS);
    F.tie_thresh = (2u * (unsigned int)eps) << (32 - DW_FIX_BITS);
    F.pad_ = 0;
    // exponent-biased operands (dw_half2d, DW_BIAS_MASK bit: 0 kl, 1 kd, 2 Sl, 3 Sd, 4 El, 5 Ed): offsets out of the constants.
    // magic - 2^20 is exact (multiple of 2^-20 below 2^33); rc is rounded once (|rc| < 2^20: ulp 1.2e-10).
    const int bm = DW_BIAS_MASK;
    F.rc = -DW_BIAS_OFFSET * (((bm & 1) ? F.w0 : 0.0) + ((bm & 16) ? F.w12 : 0.0) + ((bm & 4) ? F.w2 : 0.0));
    F.magic_b = F.magic - ((bm & 1) ? DW_BIAS_OFFSET : 0.0);
    F.t0_b = F.t0 - DW_BIAS_OFFSET * (((bm & 1) ? F.tk_l : 0.0) + ((bm & 2) ? F.tk_d : 0.0));
}

static void make_step_coef(const dw_config &c, double L, StepCoef &s) {
    const double a = c.adjacent_kernel[0];
    const double cL = c.S * L / c.sigma;
    const double Al0 = c.albedo_bare * c.p, A0 = c.albedo_bare * c.p * (8.0 * a);
    const double cl = (c.albedo_light - c.albedo_bare) / 1000.0, cd = (c.albedo_dark - c.albedo_bare) / 1000.0;
    const double g2 = c.g * c.g;
    s.x0 = g2 * (cL + (c.q - cL) * A0 + (c.q2 - c.q) * Al0 - c.q2 * c.albedo_light);
    s.xs_l = g2 * ((c.q - cL) * a * cl);
    s.xs_d = g2 * ((c.q - cL) * a * cd);
    s.SL = c.S * L;
    const int bm = DW_BIAS_MASK;
    const double xk_l = g2 * ((c.q2 - c.q) * cl), xk_d = g2 * ((c.q2 - c.q) * cd);       // FastCoef::xk_l, xk_d
    s.x0_b = s.x0 - DW_BIAS_OFFSET * (((bm & 1) ? xk_l : 0.0) + ((bm & 2) ? xk_d : 0.0) + ((bm & 4) ? s.xs_l : 0.0) + ((bm & 8) ? s.xs_d : 0.0));
}

static int ensure_lattice_buffers(dw_handle *h) {
    const size_t B = h->cfg.batch, NN = h->NN;
    int rc = dev_alloc(h, &h->lat[0], B * NN);
    if (!rc) rc = dev_alloc(h, &h->lat[1], B * NN);
    if (!rc) rc = dev_alloc(h, &h->lat_pre, B * NN);
    if (!rc) rc = dev_alloc(h, &h->slow_count, (size_t)2);
    return rc;
}

// bring the state onto the packed lattice; *converted = false if some cover is not an exact k/1000
static int grid_to_lattice(dw_handle *h, bool *converted) {
    const size_t B = h->cfg.batch, NN = h->NN;
    int rc = ensure_lattice_buffers(h);
    if (rc) return rc;
    DW_CUDA_TRY(h, cudaMemsetAsync(h->slow_count + 1, 0, sizeof(unsigned int), h->stream));
    k_grid_to_lattice<<<grid_for(B * NN), 256, 0, h->stream>>>((int)B, NN, h->grid[h->cur], h->lat[h->lcur], h->slow_count + 1);
    DW_LAUNCHED(h);
    unsigned int off = 0;
    DW_CUDA_TRY(h, cudaMemcpyAsync(&off, h->slow_count + 1, sizeof(off), cudaMemcpyDeviceToHost, h->stream));
    DW_CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    *converted = off == 0;
    if (*converted) h->lat_valid = true;
    return DW_OK;
}

static bool same_clock(const dw_clock &a, const dw_clock &b) {
    return a.L == b.L && a.dL == b.dL && a.min_L == b.min_L && a.max_L == b.max_L && a.ddL == b.ddL && a.step_count == b.step_count &&
           a.ramp_period == b.ramp_period && a.ramp_up_down == b.ramp_up_down;
}

static int launch_fused(dw_handle *h, int K, int policy, const int8_t *act_dev, uint64_t seed, unsigned int *alive) {
    if (K > DW_FUSED_MAX_STEPS) return dw_fail(h, DW_E_INVALID, "launch_fused", "K too large");
    FusedArgs A{};
    A.P = make_params(h);
    make_fast_coef(h->cfg, A.F);
    // Per-step coefficient table (luminosity coefficients + resolved policy). Callers that advance one step per launch (the
    // MLP policy, the ES population rollout) would rebuild and re-upload it every step: the table is built with look-ahead
    // and kept on the device while the clock, the constants and the policy stay on the same track.
    int rc = dev_alloc(h, &h->sc_dev, (size_t)DW_FUSED_MAX_STEPS);
    if (rc) return rc;
    long long off = -1;
    if (!h->sc_clk.empty() && policy == h->sc_policy && seed == h->sc_seed && h->epsilon == h->sc_eps &&
        !memcmp(&h->cfg, &h->sc_cfg, sizeof(dw_config))) {
        const long long j = h->clk.step_count - h->sc_clk[0].step_count;
        if (j >= 0 && j + K <= (long long)h->sc_clk.size() - 1 && same_clock(h->sc_clk[j], h->clk)) off = j;
    }
    if (off < 0) {
        const int build = K >= 64 ? K : 512;
        std::vector<StepCoef> table(build);
        h->sc_clk.assign(build + 1, h->clk);
        dw_clock c = h->clk;
        for (int j = 0; j < build; ++j) {
            h->sc_clk[j] = c;
            make_step_coef(h->cfg, c.L, table[j]);
            table[j].policy = dw_resolve_policy(policy, h->epsilon, seed, (uint32_t)(h->clk.step_count + j));
            table[j].pad_ = 0;
            update_L(c);
        }
        h->sc_clk[build] = c;
        h->sc_policy = policy; h->sc_seed = seed; h->sc_eps = h->epsilon; h->sc_cfg = h->cfg;
        // pageable source: cudaMemcpyAsync returns once the data is staged, so the local table may go out of scope
        DW_CUDA_TRY(h, cudaMemcpyAsync(h->sc_dev, table.data(), build * sizeof(StepCoef), cudaMemcpyHostToDevice, h->stream));
        off = 0;
    }
    const double L_last = h->sc_clk[off + K - 1].L;
    const dw_clock clk = h->sc_clk[off + K];
    A.sc = h->sc_dev + off;
    A.lat_in = h->lat[h->lcur];
    A.lat_out = h->lat[1 - h->lcur];
    A.lat_pre = h->lat_pre;
    A.agent_xy = h->agent_xy; A.agent_state = h->agent_state; A.actions = act_dev;
    A.done_at = h->done_at; A.agents_done_at = h->agents_done_at; A.alive = alive;
    A.reward = h->reward; A.done = h->done;
    A.seed = seed; A.step0 = (unsigned int)h->clk.step_count; A.world0 = h->world0;
    A.K = K; A.policy = policy;
    A.count_life = h->count_life ? 1 : 0;
    A.slow_count = h->slow_count;
    A.alive_mask = h->mask_on ? h->alive_mask : nullptr;
    // kernel selection: 64x64 worlds with <= 32 agents run the persistent kernel (dynamic work queue); DW_FUSED_IMPL
    // overrides for experiments: "persist" (default) | "simple" (one CTA per world) | "generic" (any N)
    const char *impl = getenv("DW_FUSED_IMPL");
    const bool n64 = h->cfg.dim == 64 && !(impl && !strcmp(impl, "generic"));
    const bool persist = n64 && !(impl && !strcmp(impl, "simple")) && h->cfg.n_agents <= DW_N64_MAX_AGENTS;
    const int dimN = h->cfg.dim;
    const bool sub64 = !(impl && !strcmp(impl, "generic")) && (dimN == 8 || dimN == 16 || dimN == 32) &&
                       (64 / dimN) * (64 / dimN) * h->cfg.n_agents <= DW_SUB64_MAX_AGENTS;
    if (policy == DW_POLICY_MLP) {           // in-kernel policy: the persistent kernels only (callers check mlp_fusable)
        if (!(persist || sub64) || h->series_on || !h->mlp_set || h->pre != PRE_LAT)
            return dw_fail(h, DW_E_UNSUPPORTED, "launch_fused", "DW_POLICY_MLP is fused only in the persistent kernels (64x64, 8x8, 16x16, 32x32)");
        A.rew_series = h->pop_rew_on ? h->pop_rew : nullptr;
        A.mlp_w = h->mlp_dev;
        A.mlp_wpm = h->pop_members ? h->cfg.batch / h->pop_members : 0;
        A.mlp_half = h->cfg.n_agents / 2;
        A.mlp_adv = h->pop_adversary;
        A.SL_prev = h->cfg.S * h->L_last;
    }
    if (h->profiling) DW_CUDA_TRY(h, cudaEventRecord(h->ev[0], h->stream));
    if (persist) {
        if (!h->persist_blocks) {
            int per_sm = 0, sms = 0;
            DW_CUDA_TRY(h, cudaFuncSetAttribute(k_fused_n64_persist<false>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                                cudaSharedmemCarveoutMaxShared));
            DW_CUDA_TRY(h, cudaFuncSetAttribute(k_fused_n64_persist<true>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                                cudaSharedmemCarveoutMaxShared));
            DW_CUDA_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_fused_n64_persist<true>, 256, 0));
            DW_CUDA_TRY(h, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->cfg.device));
            if (per_sm < 1) return dw_fail(h, DW_E_UNSUPPORTED, "launch_fused", "persistent kernel does not fit on an SM");
            h->persist_blocks = per_sm * sms;
        }
        const char *kc_env = getenv("DW_PERSIST_KC");
        A.Kc = kc_env ? atoi(kc_env) : 16;
        if (A.Kc < 1) A.Kc = 1;
        A.n_pairs = h->cfg.batch;
        A.n_chunks = (K + A.Kc - 1) / A.Kc;
        if (A.n_chunks > 1) {
            rc = dev_alloc(h, &h->persist_sync, (size_t)h->cfg.batch + 1);
            if (rc) return rc;
            DW_CUDA_TRY(h, cudaMemsetAsync(h->persist_sync, 0, ((size_t)A.n_pairs + 1) * sizeof(unsigned int), h->stream));
            A.queue = h->persist_sync;
            A.pair_done = h->persist_sync + 1;
        }                                             // else: one chunk per world (step()): static assignment, no queue to zero
        A.lat = h->lat[h->lcur];                      // in place
        const long long items = (long long)A.n_pairs * A.n_chunks;
        const int grid = (int)(items < h->persist_blocks ? items : h->persist_blocks);
        if (policy == DW_POLICY_MLP) {
            if (!h->persist_blocks_mlp) {
                int per_sm = 0, sms = 0;
                DW_CUDA_TRY(h, cudaFuncSetAttribute(k_fused_n64_persist<false, true>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                                    cudaSharedmemCarveoutMaxShared));
                DW_CUDA_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_fused_n64_persist<false, true>, 256, 0));
                DW_CUDA_TRY(h, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->cfg.device));
                if (per_sm < 1) return dw_fail(h, DW_E_UNSUPPORTED, "launch_fused", "persistent MLP kernel does not fit on an SM");
                h->persist_blocks_mlp = per_sm * sms;
            }
            const int gm = (int)(items < h->persist_blocks_mlp ? items : h->persist_blocks_mlp);
            k_fused_n64_persist<false, true><<<gm, 256, 0, h->stream>>>(A);
        } else if (h->series_on) {
            A.series_T = h->series_T + h->series_pos;
            A.series_l = h->series_l + h->series_pos;
            A.series_d = h->series_d + h->series_pos;
            k_fused_n64_persist<true><<<grid, 256, 0, h->stream>>>(A);
            h->series_pos += K;
        } else {
            k_fused_n64_persist<false><<<grid, 256, 0, h->stream>>>(A);
        }
    } else if (sub64) {
        // worlds of 8x8, 16x16 or 32x32: (64/N)^2 of them per CTA in the 4x4-tile kernel, same persistent queue
        const int N = h->cfg.dim, W = (64 / N) * (64 / N);
        void (*kern)(const FusedArgs) = N == 8 ? k_fused_sub64_persist<8> : (N == 16 ? k_fused_sub64_persist<16> : k_fused_sub64_persist<32>);
        const bool mlp = policy == DW_POLICY_MLP;
        const size_t dyn = mlp ? DW_SUB64_MLP_SMEM : 0;
        if (mlp) kern = N == 8 ? k_fused_sub64_persist<8, false, true> : (N == 16 ? k_fused_sub64_persist<16, false, true> : k_fused_sub64_persist<32, false, true>);
        if (h->series_on) {
            kern = N == 8 ? k_fused_sub64_persist<8, true> : (N == 16 ? k_fused_sub64_persist<16, true> : k_fused_sub64_persist<32, true>);
            A.series_T = h->series_T + h->series_pos;
            A.series_l = h->series_l + h->series_pos;
            A.series_d = h->series_d + h->series_pos;
            h->series_pos += K;
        }
        int &blocks = mlp ? h->sub64_blocks_mlp : (h->series_on ? h->sub64_blocks_series : h->sub64_blocks);
        if (!blocks) {
            int per_sm = 0, sms = 0;
            if (dyn) DW_CUDA_TRY(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
            DW_CUDA_TRY(h, cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            DW_CUDA_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, dyn));
            DW_CUDA_TRY(h, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->cfg.device));
            if (per_sm < 1) return dw_fail(h, DW_E_UNSUPPORTED, "launch_fused", "sub-64 kernel does not fit on an SM");
            blocks = per_sm * sms;
        }
        const char *kc_env = getenv("DW_PERSIST_KC");
        A.Kc = kc_env ? atoi(kc_env) : 16;
        if (A.Kc < 1) A.Kc = 1;
        A.n_pairs = (h->cfg.batch + W - 1) / W;
        A.n_chunks = (K + A.Kc - 1) / A.Kc;
        if (A.n_chunks > 1) {
            rc = dev_alloc(h, &h->persist_sync, (size_t)h->cfg.batch + 1);
            if (rc) return rc;
            DW_CUDA_TRY(h, cudaMemsetAsync(h->persist_sync, 0, ((size_t)A.n_pairs + 1) * sizeof(unsigned int), h->stream));
            A.queue = h->persist_sync;
            A.pair_done = h->persist_sync + 1;
        }
        A.lat = h->lat[h->lcur];                      // in place
        const long long items = (long long)A.n_pairs * A.n_chunks;
        const int grid = (int)(items < blocks ? items : blocks);
        kern<<<grid, 256, dyn, h->stream>>>(A);
    } else {
        size_t smem = fused_smem_bytes(h->cfg.dim, h->cfg.n_agents);
        // 4x4-tile kernel for every side >= 12 that fits with rows padded to a multiple of 4 words
        const size_t smem_t4 = smem + 2 * (size_t)dimN * (size_t)(((dimN + 3) & ~3) - dimN) * sizeof(uint32_t);
        const bool tile4_fits = smem_t4 <= 200 * 1024;
        if (smem > 48 * 1024 && !h->fused_attr_set) {      // both one-CTA-per-world fallbacks (64x64 with > 817 agents needs it too)
            DW_CUDA_TRY(h, cudaFuncSetAttribute(k_fused_generic, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            DW_CUDA_TRY(h, cudaFuncSetAttribute(k_fused_n64, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            h->fused_attr_set = true;
        }
        const bool tile4 = !n64 && dimN >= 12 && tile4_fits && !(impl && !strcmp(impl, "generic")) && !getenv("DW_TILE4_MULT4_ONLY") ||
                           (!n64 && dimN % 4 == 0 && dimN >= 12 && !(impl && !strcmp(impl, "generic")));
        if (tile4) smem = smem_t4;
        if (n64) k_fused_n64<<<h->cfg.batch, 256, smem, h->stream>>>(A);
        else if (tile4) {
            // 4x4 tiles dealt round-robin to the threads: pick the block size that keeps the most useful warps resident
            // (64 registers per thread: at most 1024 threads per SM; CTAs per SM bounded by the shared-memory footprint)
            if (!h->tile4_threads) {
                DW_CUDA_TRY(h, cudaFuncSetAttribute(k_fused_tile4<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                DW_CUDA_TRY(h, cudaFuncSetAttribute(k_fused_tile4<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                const int TT = ((dimN + 3) / 4) * ((dimN + 3) / 4);
                const int cps_max = (int)std::min<size_t>(16, (227 * 1024) / (smem + 1024));
                double best = -1.0;
                for (int cps = 1; cps <= std::max(1, cps_max); ++cps) {
                    const int cap = (1024 / cps) & ~31;                       // threads per CTA that still fit cps CTAs in the register file
                    if (cap < 32) break;
                    const int rounds = (TT + cap - 1) / cap;
                    const int threads = std::min(cap, (((TT + rounds - 1) / rounds) + 31) & ~31);
                    const double useful = (double)TT / ((double)rounds * threads) * std::min(1024, cps * threads);
                    if (useful > best) { best = useful; h->tile4_threads = threads; }
                }
            }
            const char *tenv = getenv("DW_TILE4_THREADS");
            const int threads = tenv ? atoi(tenv) : h->tile4_threads;
            void (*kern)(const FusedArgs) = dimN % 4 == 0 ? k_fused_tile4<false> : k_fused_tile4<true>;
            if (h->series_on) {
                if (dimN % 4 != 0) return dw_fail(h, DW_E_UNSUPPORTED, "launch_fused", "series mode of the 4x4-tile kernel: sides that are multiples of 4");
                kern = k_fused_tile4<false, true>;
                DW_CUDA_TRY(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                A.series_T = h->series_T + h->series_pos;
                A.series_l = h->series_l + h->series_pos;
                A.series_d = h->series_d + h->series_pos;
                h->series_pos += K;
            }
            if (!h->tile4_blocks || h->tile4_blocks_threads != threads || h->series_on) {
                int per_sm = 0, sms = 0;
                DW_CUDA_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
                DW_CUDA_TRY(h, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->cfg.device));
                if (per_sm < 1) return dw_fail(h, DW_E_UNSUPPORTED, "launch_fused", "4x4-tile kernel does not fit on an SM");
                h->tile4_blocks = per_sm * sms;
                h->tile4_blocks_threads = h->series_on ? 0 : threads;       // the series variant is re-queried every time
            }
            // persistent CTAs, (world, chunk) work queue, state advanced in place (like the 64x64 kernel)
            const char *kc_env = getenv("DW_PERSIST_KC");
            A.Kc = kc_env ? atoi(kc_env) : 16;
            if (A.Kc < 1) A.Kc = 1;
            A.n_pairs = h->cfg.batch;
            A.n_chunks = (K + A.Kc - 1) / A.Kc;
            if (A.n_chunks > 1) {
                rc = dev_alloc(h, &h->persist_sync, (size_t)h->cfg.batch + 1);
                if (rc) return rc;
                DW_CUDA_TRY(h, cudaMemsetAsync(h->persist_sync, 0, ((size_t)A.n_pairs + 1) * sizeof(unsigned int), h->stream));
                A.queue = h->persist_sync;
                A.pair_done = h->persist_sync + 1;
            }
            A.lat = h->lat[h->lcur];
            const long long items = (long long)A.n_pairs * A.n_chunks;
            const int grid = (int)(items < h->tile4_blocks ? items : h->tile4_blocks);
            kern<<<grid, threads, smem, h->stream>>>(A);
        } else {
            k_fused_generic<<<h->cfg.batch, 256, smem, h->stream>>>(A);
            h->lcur = 1 - h->lcur;
        }
        if (n64) h->lcur = 1 - h->lcur;
    }
    DW_LAUNCHED(h);
    if (h->profiling) {
        DW_CUDA_TRY(h, cudaEventRecord(h->ev[1], h->stream));
        h->ev_pending = true;
        h->ev_cells = (uint64_t)h->cfg.batch * h->NN * (uint64_t)K;
    }
    h->lat_valid = true;
    h->grid_valid = false;
    h->pre = PRE_LAT;
    h->L_last = L_last;
    state_changed(h);
    h->clk = clk;
    return DW_OK;
}

static int run_steps_generic(dw_handle *h, int K, int policy, const int8_t *act_dev, uint64_t seed, unsigned int *alive);

// First step of a run from the lean reset state (fp64 cover planes, off the lattice): update_agents on the planes, one
// literal forward whose new covers go straight onto the packed lattice, reward/done and lifespan counters. No 7-channel
// grid is touched; the post-graze planes stay behind as the pre-state for lazy materialisation / diagnostics.
static int lean_first_step(dw_handle *h, int policy, const int8_t *act_dev, uint64_t seed, unsigned int *alive_slot) {
    int rc = ensure_lattice_buffers(h);
    if (rc) return rc;
    if (policy == DW_POLICY_REPLAY) rc = launch_agents(h, act_dev, h->cfg.batch, h->cfg.n_agents, policy, seed, true);
    else rc = launch_agents(h, nullptr, 0, 0, policy, seed, true);
    if (rc) return rc;
    const DevParams P = make_params(h);
    DW_CUDA_TRY(h, cudaMemsetAsync(h->world_max, 0, (size_t)P.B * 2 * sizeof(unsigned long long), h->stream));
    SrcCov src{h->cov, h->NN};
    launch_forward_lattice(h, P, h->cfg.S * h->clk.L, src, h->lat[h->lcur], h->world_max);
    DW_LAUNCHED(h);
    rc = launch_stamp(h, nullptr, h->count_life, alive_slot, true);
    if (rc) return rc;
    h->pre = PRE_COV;
    h->L_last = h->clk.L;
    h->lat_valid = true;
    h->cov_valid = false;
    h->grid_valid = false;
    state_changed(h);
    update_L(h->clk);
    return DW_OK;
}

// K steps, fused where the state allows it
static int run_steps_fused(dw_handle *h, int K, int policy, const int8_t *act_dev, uint64_t seed, unsigned int *alive) {
    const size_t per_step = (size_t)h->cfg.batch * h->cfg.n_agents;
    if (!h->lat_valid && h->cov_valid) {
        int rc = lean_first_step(h, policy, act_dev, seed, alive);
        if (rc) return rc;
        K -= 1; alive += 1;
        if (act_dev) act_dev += per_step;
        if (K == 0) return DW_OK;
    }
    if (!h->lat_valid) {
        if (!h->grid_valid) return dw_fail(h, DW_E_STATE, "dw_run", "no state uploaded");
        bool ok = false;
        int rc = grid_to_lattice(h, &ok);
        if (rc) return rc;
        if (!ok) {
            // off-lattice covers (the unrounded state right after reset()): the first step must be literal
            rc = run_steps_generic(h, 1, policy, act_dev, seed, alive);
            if (rc) return rc;
            K -= 1; alive += 1;
            if (act_dev) act_dev += per_step;
            if (K == 0) return DW_OK;
            rc = grid_to_lattice(h, &ok);
            if (rc) return rc;
            if (!ok) return dw_fail(h, DW_E_STATE, "dw_run", "state is off the 0.001 lattice after a forward step");
        }
    }
    return launch_fused(h, K, policy, act_dev, seed, alive);
}

// DW_POLICY_MLP inside the fused kernels: 64x64 worlds with <= 32 agents, or 8x8 / 16x16 / 32x32 worlds, whose state is on the
// lattice with the post-graze lattice of the last step at hand (the observation windows of the first fused step come from it)
static bool mlp_fusable(const dw_handle *h) {
    if (getenv("DW_MLP_UNFUSED")) return false;
    const char *impl = getenv("DW_FUSED_IMPL");
    const int N = h->cfg.dim, n = h->cfg.n_agents;
    // The sub-64 kernel has the in-kernel policy too (k_fused_sub64_persist<N, false, true>, four agents per warp pass), but
    // several worlds share a CTA there and their agents' networks run one after the other on 8 warps: measured on the ES
    // shape (64 members x 32 worlds of 16x16, 4 agents) 37.8 us per step against 42 us for the whole per-step sequence whose
    // k_obs_mlp spreads the 8192 agents over every SM -- and 22.8 ms vs 19.8 ms per generation. It stays opt-in
    // (DW_MLP_FUSE_SUB64=1; the tests run both paths).
    const bool sub = (N == 8 || N == 16 || N == 32) && (64 / N) * (64 / N) * n <= DW_SUB64_MAX_AGENTS && getenv("DW_MLP_FUSE_SUB64");
    const bool shape = (N == 64 && n <= DW_N64_MAX_AGENTS) || sub;
    return shape && n > 0 && dw_fused_supported(h) && h->lat_valid && h->pre == PRE_LAT && h->mlp_set && !h->series_on && !impl;
}

static int stage_actions8(dw_handle *h, const int8_t *actions, size_t count) {
    if (h->action_cap < count) {
        if (h->action_dev) cudaFree(h->action_dev);
        h->action_dev = nullptr;
        DW_CUDA_TRY(h, cudaMalloc((void **)&h->action_dev, count));
        h->action_cap = count;
    }
    DW_CUDA_TRY(h, cudaMemcpyAsync(h->action_dev, actions, count, cudaMemcpyHostToDevice, h->stream));
    return DW_OK;
}

// K <= 64 steps through the materialising kernels (any N, any kernels, off-lattice states)
static int run_steps_generic(dw_handle *h, int K, int policy, const int8_t *act_dev, uint64_t seed, unsigned int *alive) {
    int rc = ensure_grid(h);
    if (rc) return rc;
    const size_t per_step = (size_t)h->cfg.batch * h->cfg.n_agents;
    for (int j = 0; j < K; ++j) {
        if (policy == DW_POLICY_REPLAY) rc = launch_agents(h, act_dev + (size_t)j * per_step, h->cfg.batch, h->cfg.n_agents, policy, seed);
        else rc = launch_agents(h, nullptr, 0, 0, policy, seed);
        if (rc) return rc;
        rc = launch_forward_tail(h, h->count_life, alive + j);
        if (rc) return rc;
    }
    return DW_OK;
}

static int run_chunk_impl(dw_handle *h, int K, int policy, const int8_t *act_dev, uint64_t seed, uint64_t *done_mask,
                          unsigned int *alive_last, int *first_all_done = nullptr) {
    if (K < 1 || K > DW_FUSED_MAX_STEPS) return dw_fail(h, DW_E_INVALID, "run_chunk", "1 <= K <= 4096");
    DW_CUDA_TRY(h, cudaMemsetAsync(h->alive, 0, DW_FUSED_MAX_STEPS * sizeof(unsigned int), h->stream));
    int rc = DW_OK;
    if (policy == DW_POLICY_MLP) {
        // 64x64 worlds on the lattice: windows + network inside the persistent kernel, all remaining steps in one launch
        // (mlp_fusable). Elsewhere the policy runs between steps on the device (observation windows from the last pre-state,
        // then the network) and each step is a one-step launch replaying the device-resident actions.
        for (int j = 0; j < K && !rc; ++j) {
            if (mlp_fusable(h)) {
                rc = launch_fused(h, K - j, DW_POLICY_MLP, nullptr, seed, h->alive + j);
                break;
            }
            rc = mlp_actions(h);
            if (rc) break;
            rc = dw_fused_supported(h) ? run_steps_fused(h, 1, DW_POLICY_REPLAY, h->action_dev, seed, h->alive + j)
                                       : run_steps_generic(h, 1, DW_POLICY_REPLAY, h->action_dev, seed, h->alive + j);
        }
    } else {
        rc = dw_fused_supported(h) ? run_steps_fused(h, K, policy, act_dev, seed, h->alive)
                                   : run_steps_generic(h, K, policy, act_dev, seed, h->alive);
    }
    if (rc) return rc;
    std::vector<unsigned int> alive(K);
    DW_CUDA_TRY(h, cudaMemcpyAsync(alive.data(), h->alive, K * sizeof(unsigned int), cudaMemcpyDeviceToHost, h->stream));
    DW_CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    if (h->ev_pending) {
        float ms = 0.f;
        DW_CUDA_TRY(h, cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]));
        h->prof.fused_launches += 1;
        h->prof.fused_ms += ms;
        h->prof.fused_cell_updates += h->ev_cells;
        h->ev_pending = false;
    }
    uint64_t m = 0;
    int first = -1;
    for (int j = 0; j < K; ++j) if (alive[j] == 0) { if (j < 64) m |= (1ull << j); if (first < 0) first = j; }
    if (done_mask) *done_mask = m;
    if (first_all_done) *first_all_done = first;
    if (alive_last) *alive_last = alive[K - 1];
    return DW_OK;
}

static int check_policy(dw_handle *h, int policy, const int8_t *actions) {
    if (h->agents_open) return dw_fail(h, DW_E_STATE, "dw_run", "dw_agents_begin not closed by dw_agents_collide");
    if (policy < 0 || policy > DW_POLICY_MLP) return dw_fail(h, DW_E_INVALID, "policy", "unknown policy");
    if (policy == DW_POLICY_MLP && h->cfg.n_agents > 0 && !h->mlp_set) return dw_fail(h, DW_E_STATE, "policy", "DW_POLICY_MLP needs dw_set_mlp");
    if (policy == DW_POLICY_REPLAY && h->cfg.n_agents > 0 && !actions) return dw_fail(h, DW_E_INVALID, "policy", "REPLAY needs actions[K,B,n]");
    return DW_OK;
}

extern "C" int dw_run_chunk(dw_handle *h, int32_t K, int32_t policy, const int8_t *actions, uint64_t seed, uint64_t *done_mask) {
    if (!h || K < 1 || K > 64) return DW_E_INVALID;
    DW_CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    int rc = check_policy(h, policy, actions);
    if (rc) return rc;
    const size_t per_step = (size_t)h->cfg.batch * h->cfg.n_agents;
    if (policy == DW_POLICY_REPLAY && per_step) {
        rc = stage_actions8(h, actions, per_step * K);
        if (rc) return rc;
    }
    return run_chunk_impl(h, K, policy, h->action_dev, seed, done_mask, nullptr);
}

// Statistics-only lifespan runs (ensemble.simulate_lifespan): the notebook's loop stops at the first step at which every world
// is grid_done, which is only known after a chunk has run. dw_run(stop_all_done) rewinds to a checkpoint and replays up to that
// step (exact final state); a caller that only wants the lifespan counters can instead let the chunk run on and take the
// surplus out of agents_done_at afterwards -- done_at cannot grow after the stopping step (every world is done for good), and
// the kernels record per agent which steps of the chunk it was not done in. No checkpoint copy, no replay.
static bool trim_supported(const dw_handle *h, int policy) {
    const int N = h->cfg.dim, n = h->cfg.n_agents;
    const bool fam = (N == 64 && n <= DW_N64_MAX_AGENTS) || ((N == 8 || N == 16 || N == 32) && (64 / N) * (64 / N) * n <= DW_SUB64_MAX_AGENTS);
    return fam && h->lat_valid && dw_fused_supported(h) && policy != DW_POLICY_MLP && !getenv("DW_FUSED_IMPL") && !getenv("DW_NO_TRIM");
}

extern "C" int dw_trim_supported(dw_handle *h, int32_t policy, int32_t *yes) {
    if (!h || !yes) return DW_E_INVALID;
    *yes = trim_supported(h, policy) ? 1 : 0;
    return DW_OK;
}

extern "C" int dw_run_chunk_masked(dw_handle *h, int32_t K, int32_t policy, const int8_t *actions, uint64_t seed, uint64_t *done_mask) {
    if (!h || K < 1 || K > 64) return DW_E_INVALID;
    DW_CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    int rc = check_policy(h, policy, actions);
    if (rc) return rc;
    if (!trim_supported(h, policy)) return dw_fail(h, DW_E_STATE, "dw_run_chunk_masked", "ask dw_trim_supported first");
    const size_t per_step = (size_t)h->cfg.batch * h->cfg.n_agents;
    if (policy == DW_POLICY_REPLAY && per_step) {
        rc = stage_actions8(h, actions, per_step * K);
        if (rc) return rc;
    }
    rc = dev_alloc(h, &h->alive_mask, per_step);
    if (rc) return rc;
    DW_CUDA_TRY(h, cudaMemsetAsync(h->alive_mask, 0, (per_step ? per_step : 1) * sizeof(unsigned long long), h->stream));
    h->mask_on = true;
    rc = run_chunk_impl(h, K, policy, h->action_dev, seed, done_mask, nullptr);
    h->mask_on = false;
    h->mask_complete = rc == DW_OK;
    return rc;
}

extern "C" int dw_trim_lifespans(dw_handle *h, int32_t j) {
    if (!h || j < 0 || j > 63) return DW_E_INVALID;
    DW_CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    if (!h->mask_complete) return dw_fail(h, DW_E_STATE, "dw_trim_lifespans", "no masked chunk to trim (dw_run_chunk_masked)");
    const size_t count = (size_t)h->cfg.batch * h->cfg.n_agents;
    if (count) {
        k_trim_lifespans<<<(unsigned)((count + 255) / 256), 256, 0, h->stream>>>(count, j, h->alive_mask, h->agents_done_at);
        DW_LAUNCHED(h);
    }
    h->mask_complete = false;
    return DW_OK;
}

extern "C" int dw_run(dw_handle *h, int64_t K, int32_t policy, const int8_t *actions, uint64_t seed, int32_t stop_all_done,
                      dw_run_result *res) {
    if (!h || K < 0) return DW_E_INVALID;
    DW_CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    int rc = check_policy(h, policy, actions);
    if (rc) return rc;
    const size_t per_step = (size_t)h->cfg.batch * h->cfg.n_agents;
    int64_t done_steps = 0;
    unsigned int alive_last = (unsigned int)h->cfg.batch;
    int hit = 0;
    // Segment length: without the stopping rule one launch covers up to DW_FUSED_MAX_STEPS steps; with it, segments
    // of 64 steps bound the work that has to be replayed when the all-done step falls inside a segment.
    const int64_t seg = stop_all_done ? 64 : DW_FUSED_MAX_STEPS;
    while (done_steps < K) {
        const int k = (int)((K - done_steps) < seg ? (K - done_steps) : seg);
        if (policy == DW_POLICY_REPLAY && per_step) {
            rc = stage_actions8(h, actions + (size_t)done_steps * per_step, per_step * k);
            if (rc) return rc;
        }
        if (stop_all_done) {
            rc = ckpt_save(h, 1);
            if (rc) return rc;
        }
        int first = -1;
        rc = run_chunk_impl(h, k, policy, h->action_dev, seed, nullptr, &alive_last, &first);
        if (rc) return rc;
        if (stop_all_done && first >= 0) {
            if (first < k - 1) {     // overshot the notebook's stopping step: rewind and replay exactly first+1 steps
                rc = ckpt_restore(h, 1);
                if (rc) return rc;
                rc = run_chunk_impl(h, first + 1, policy, h->action_dev, seed, nullptr, &alive_last, nullptr);
                if (rc) return rc;
            }
            done_steps += first + 1;
            hit = 1;
            break;
        }
        done_steps += k;
    }
    if (res) {
        res->steps_run = done_steps;
        res->worlds_alive = alive_last;
        res->all_done_hit = hit;
        res->_pad = 0;
    }
    return DW_OK;
}

// ---- ES fitness rollout of a whole population (daisy/evo/sges.py:144-181) -------------------------------------------
extern "C" int dw_run_population(dw_handle *h, int64_t max_steps, int64_t *steps_run) {
    if (!h || max_steps < 0) return DW_E_INVALID;
    DW_CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    if (!h->pop_members || !h->mlp_set) return dw_fail(h, DW_E_STATE, "dw_run_population", "call dw_set_mlp_population first");
    const int P = h->pop_members, n = h->cfg.n_agents, wpm = h->cfg.batch / P;
    if (n < 2) return dw_fail(h, DW_E_INVALID, "dw_run_population", "needs at least two agents per world (member half + adversary half)");
    const size_t Bn = (size_t)h->cfg.batch * n;
    int rc = dev_alloc(h, &h->pop_sum, (size_t)P);
    if (!rc) rc = dev_alloc(h, &h->pop_done, (size_t)P);
    if (!rc) rc = dev_alloc(h, &h->pop_steps, (size_t)P);
    if (!rc) rc = dev_alloc(h, &h->pop_frozen, Bn);
    if (!rc) rc = dev_alloc(h, &h->pop_ndone, (size_t)1);
    if (rc) return rc;
    DW_CUDA_TRY(h, cudaMemsetAsync(h->pop_sum, 0, P * sizeof(double), h->stream));
    DW_CUDA_TRY(h, cudaMemsetAsync(h->pop_done, 0, P * sizeof(int), h->stream));
    DW_CUDA_TRY(h, cudaMemsetAsync(h->pop_steps, 0, P * sizeof(int64_t), h->stream));
    DW_CUDA_TRY(h, cudaMemsetAsync(h->pop_frozen, 0, Bn * sizeof(int64_t), h->stream));
    DW_CUDA_TRY(h, cudaMemsetAsync(h->pop_ndone, 0, sizeof(unsigned int), h->stream));
    rc = dw_reset_lifespans(h);
    if (rc) return rc;
    int64_t t = 0;
    while (t < max_steps) {
        if (mlp_fusable(h) && !getenv("DW_POP_UNFUSED")) {
            // a segment of up to 64 steps in ONE launch (policy inside the kernel, per-step agent states recorded), then the
            // members' bookkeeping of those steps in one post-pass; one small read-back per segment
            const int S = (int)std::min<int64_t>(64, max_steps - t);
            rc = dev_alloc(h, &h->pop_rew, (size_t)64 * Bn);
            if (rc) return rc;
            DW_CUDA_TRY(h, cudaMemsetAsync(h->alive, 0, DW_FUSED_MAX_STEPS * sizeof(unsigned int), h->stream));
            const long long step0 = (long long)h->clk.step_count;
            h->pop_rew_on = true;
            rc = launch_fused(h, S, DW_POLICY_MLP, nullptr, 0, h->alive);
            h->pop_rew_on = false;
            if (rc) return rc;
            k_pop_post<<<P, 128, 0, h->stream>>>(S, wpm, n, n / 2, h->pop_rew, Bn, h->pop_sum, h->pop_done, h->pop_steps, h->pop_frozen, step0,
                                                 h->pop_ndone);
            DW_LAUNCHED(h);
            t += S;
            unsigned int nd = 0;
            DW_CUDA_TRY(h, cudaMemcpyAsync(&nd, h->pop_ndone, sizeof(nd), cudaMemcpyDeviceToHost, h->stream));
            DW_CUDA_TRY(h, cudaStreamSynchronize(h->stream));
            if ((int)nd == P) break;
            continue;
        }
        DW_CUDA_TRY(h, cudaMemsetAsync(h->alive, 0, sizeof(unsigned int), h->stream));
        rc = mlp_actions(h);
        if (rc) return rc;
        rc = dw_fused_supported(h) ? run_steps_fused(h, 1, DW_POLICY_REPLAY, h->action_dev, 0, h->alive)
                                   : run_steps_generic(h, 1, DW_POLICY_REPLAY, h->action_dev, 0, h->alive);
        if (rc) return rc;
        t += 1;
        k_pop_accumulate<<<P, 128, 0, h->stream>>>(wpm, n, n / 2, h->reward, h->done, h->agents_done_at, h->pop_sum, h->pop_done,
                                                   h->pop_steps, h->pop_frozen, (long long)h->clk.step_count, h->pop_ndone);
        DW_LAUNCHED(h);
        if ((t & 7) == 0 || t == max_steps) {            // every member's loop over? (one small read-back per 8 steps)
            unsigned int nd = 0;
            DW_CUDA_TRY(h, cudaMemcpyAsync(&nd, h->pop_ndone, sizeof(nd), cudaMemcpyDeviceToHost, h->stream));
            DW_CUDA_TRY(h, cudaStreamSynchronize(h->stream));
            if ((int)nd == P) break;
        }
    }
    if (steps_run) *steps_run = t;
    return DW_OK;
}

extern "C" int dw_get_population_results(dw_handle *h, double *fitness, int64_t *member_steps, int64_t *total_steps) {
    if (!h) return DW_E_INVALID;
    DW_CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    if (!h->pop_members || !h->pop_sum) return dw_fail(h, DW_E_STATE, "dw_get_population_results", "no population rollout has run");
    const int P = h->pop_members, n = h->cfg.n_agents, wpm = h->cfg.batch / P;
    std::vector<double> sum(P);
    DW_CUDA_TRY(h, cudaMemcpyAsync(sum.data(), h->pop_sum, P * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (member_steps) DW_CUDA_TRY(h, cudaMemcpyAsync(member_steps, h->pop_steps, P * sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
    if (total_steps)
        DW_CUDA_TRY(h, cudaMemcpyAsync(total_steps, h->pop_frozen, (size_t)h->cfg.batch * n * sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
    DW_CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    if (fitness) for (int m = 0; m < P; ++m) fitness[m] = sum[m] / (double)(wpm * n);     // sges.py:179
    return DW_OK;
}

// ---- per-step ensemble series from inside the fused kernel ---------------------------------------------------------------
// K steps with the on-device policy; out[K][3] = per step {global mean of the unrounded temperature of that step's forward
// (env.temp.mean(), notebook_helpers.py:50), mean light cover, mean dark cover after the step}, reduced in the kernel
// (warp shuffles -> per-world shared atomics -> one global atomic per world-step). 64x64 worlds with <= 32 agents.
extern "C" int dw_run_series(dw_handle *h, int64_t K, int32_t policy, const int8_t *actions, uint64_t seed, double *out) {
    if (!h || !out || K < 1 || K > DW_FUSED_MAX_STEPS) return DW_E_INVALID;
    DW_CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    const int dN = h->cfg.dim;
    const bool n64_ok = dN == 64 && h->cfg.n_agents <= DW_N64_MAX_AGENTS;
    const bool sub_ok = (dN == 8 || dN == 16 || dN == 32) && (64 / dN) * (64 / dN) * h->cfg.n_agents <= DW_SUB64_MAX_AGENTS;
    const bool t4_ok = dN != 64 && !sub_ok && dN >= 20 && dN % 4 == 0;     // k_fused_tile4<false, true>
    if (!(n64_ok || sub_ok || t4_ok) || !dw_fused_supported(h) || policy == DW_POLICY_MLP || getenv("DW_FUSED_IMPL"))
        return dw_fail(h, DW_E_UNSUPPORTED, "dw_run_series",
                       "series mode runs in the persistent kernels (64x64 with <= 32 agents; 8x8, 16x16, 32x32 with <= 256 agents per CTA; "
                       "other multiples of 4 from 20 that fit in shared memory; built-in policies)");
    int rc = dev_alloc(h, &h->series_T, (size_t)DW_FUSED_MAX_STEPS);
    if (!rc) rc = dev_alloc(h, &h->series_l, (size_t)DW_FUSED_MAX_STEPS);
    if (!rc) rc = dev_alloc(h, &h->series_d, (size_t)DW_FUSED_MAX_STEPS);
    if (rc) return rc;
    DW_CUDA_TRY(h, cudaMemsetAsync(h->series_T, 0, K * sizeof(double), h->stream));
    DW_CUDA_TRY(h, cudaMemsetAsync(h->series_l, 0, K * sizeof(unsigned long long), h->stream));
    DW_CUDA_TRY(h, cudaMemsetAsync(h->series_d, 0, K * sizeof(unsigned long long), h->stream));
    const double cells = (double)h->cfg.batch * (double)h->NN;
    int64_t done = 0;
    if (!h->lat_valid) {
        // the step that leaves the off-lattice reset state is literal: its statistics come from the lazy diagnostics
        dw_run_result r;
        rc = dw_run(h, 1, policy, actions, seed, 0, &r);
        if (rc) return rc;
        double t[4], c[4];
        rc = dw_get_diag_stats(h, DW_DIAG_TEMP, t);
        if (!rc) rc = dw_get_cover_stats(h, c);
        if (rc) return rc;
        out[0] = t[0]; out[1] = c[0]; out[2] = c[1];
        done = 1;
        if (actions) actions += (size_t)h->cfg.batch * h->cfg.n_agents;
    }
    if (done < K) {
        h->series_on = true;
        h->series_pos = 0;
        dw_run_result r;
        rc = dw_run(h, K - done, policy, actions, seed, 0, &r);
        h->series_on = false;
        if (rc) return rc;
        const int64_t m = K - done;
        std::vector<double> T(m);
        std::vector<unsigned long long> l(m), d(m);
        DW_CUDA_TRY(h, cudaMemcpyAsync(T.data(), h->series_T, m * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        DW_CUDA_TRY(h, cudaMemcpyAsync(l.data(), h->series_l, m * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
        DW_CUDA_TRY(h, cudaMemcpyAsync(d.data(), h->series_d, m * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
        DW_CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        const double inv_sqrt_g = 1.0 / sqrt(h->cfg.g);
        for (int64_t j = 0; j < m; ++j) {
            out[3 * (done + j)] = T[j] * inv_sqrt_g / cells;
            out[3 * (done + j) + 1] = (double)l[j] / 1000.0 / cells;
            out[3 * (done + j) + 2] = (double)d[j] / 1000.0 / cells;
        }
    }
    return DW_OK;
}
