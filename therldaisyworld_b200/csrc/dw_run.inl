// dw_run / dw_run_chunk: K steps with an on-device policy and the notebook lifespan counters.
// Included at the end of dw_api.cu.

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
static int run_steps_generic(dw_handle *h, int K, int policy, const int8_t *act_dev, uint64_t seed) {
    int rc = ensure_grid(h);
    if (rc) return rc;
    const size_t per_step = (size_t)h->cfg.batch * h->cfg.n_agents;
    for (int j = 0; j < K; ++j) {
        if (policy == DW_POLICY_REPLAY) rc = launch_agents(h, act_dev + (size_t)j * per_step, h->cfg.batch, h->cfg.n_agents, policy, seed);
        else rc = launch_agents(h, nullptr, 0, 0, policy, seed);
        if (rc) return rc;
        rc = launch_forward_tail(h, true, h->alive + j);
        if (rc) return rc;
    }
    return DW_OK;
}

static int run_chunk_impl(dw_handle *h, int K, int policy, const int8_t *act_dev, uint64_t seed, uint64_t *done_mask,
                          unsigned int *alive_last) {
    if (K < 1 || K > 64) return dw_fail(h, DW_E_INVALID, "run_chunk", "1 <= K <= 64");
    DW_CUDA_TRY(h, cudaMemsetAsync(h->alive, 0, 64 * sizeof(unsigned int), h->stream));
    int rc = dw_fused_supported(h) ? run_steps_fused(h, K, policy, act_dev, seed) : run_steps_generic(h, K, policy, act_dev, seed);
    if (rc) return rc;
    unsigned int alive[64];
    DW_CUDA_TRY(h, cudaMemcpyAsync(alive, h->alive, K * sizeof(unsigned int), cudaMemcpyDeviceToHost, h->stream));
    DW_CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    uint64_t m = 0;
    for (int j = 0; j < K; ++j) if (alive[j] == 0) m |= (1ull << j);
    if (done_mask) *done_mask = m;
    if (alive_last) *alive_last = alive[K - 1];
    return DW_OK;
}

static int check_policy(dw_handle *h, int policy, const int8_t *actions) {
    if (policy < 0 || policy > DW_POLICY_RANDOM) return dw_fail(h, DW_E_INVALID, "policy", "unknown policy");
    if (policy == DW_POLICY_REPLAY && h->cfg.n_agents > 0 && !actions) return dw_fail(h, DW_E_INVALID, "policy", "REPLAY needs actions[K,B,n]");
    return DW_OK;
}

extern "C" int dw_run_chunk(dw_handle *h, int32_t K, int32_t policy, const int8_t *actions, uint64_t seed, uint64_t *done_mask) {
    if (!h) return DW_E_INVALID;
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
    while (done_steps < K) {
        const int k = (int)((K - done_steps) < 64 ? (K - done_steps) : 64);
        if (policy == DW_POLICY_REPLAY && per_step) {
            rc = stage_actions8(h, actions + (size_t)done_steps * per_step, per_step * k);
            if (rc) return rc;
        }
        if (stop_all_done) {
            rc = ckpt_save(h, 1);
            if (rc) return rc;
        }
        uint64_t mask = 0;
        rc = run_chunk_impl(h, k, policy, h->action_dev, seed, &mask, &alive_last);
        if (rc) return rc;
        if (stop_all_done && mask) {
            int j = 0;
            while (!((mask >> j) & 1ull)) ++j;
            if (j < k - 1) {     // overshot the notebook's stopping step: rewind and replay exactly j+1 steps
                rc = ckpt_restore(h, 1);
                if (rc) return rc;
                rc = run_chunk_impl(h, j + 1, policy, h->action_dev, seed, &mask, &alive_last);
                if (rc) return rc;
            }
            done_steps += j + 1;
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
