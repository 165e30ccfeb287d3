// footsies_kernels.cu -- C ABI of the batched FOOTSIES simulator (include/footsies_b200.h) on top of the kernels in
// step_kernel.cuh.
#include <stdlib.h>

#include "step_kernel.cuh"

using namespace fg;
using namespace fgk;

namespace {

thread_local char g_err[512] = "";
int fail(int code, const char *fmt, const char *detail = "") {
    snprintf(g_err, sizeof g_err, fmt, detail);
    return code;
}
#define CUDA_TRY(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) \
    return fail(FG_ERR_CUDA, #expr ": %s", cudaGetErrorString(_e)); } while (0)

}  // namespace

struct fg_handle {
    fg_config cfg;
    fg_buffers buf;
    bool bound;
    Tables *d_tables;
    int sm_count;
    int64_t launches;
    uint8_t *d_mask;       // staging for fg_reset_host
    int large_shape_min_envs;
};

namespace {

int grid_for(const fg_handle *h, int blocks_per_sm) {
    int want = (h->cfg.num_envs + kThreads - 1) / kThreads;
    int cap = h->sm_count * blocks_per_sm;
    return want < cap ? (want > 0 ? want : 1) : cap;
}

Params make_params(const fg_handle *h) {
    Params p;
    memset(&p, 0, sizeof p);
    p.pl_f1 = (uint4 *)h->buf.state[FG_PLANE_F1]; p.pl_f2 = (uint4 *)h->buf.state[FG_PLANE_F2];
    p.pl_env = (uint4 *)h->buf.state[FG_PLANE_ENV]; p.pl_rng = (uint4 *)h->buf.state[FG_PLANE_RNG];
    p.stats = (unsigned long long *)h->buf.stats;
    p.act1 = h->buf.actions_p1; p.act2 = h->buf.actions_p2;
    p.obs = (float4 *)h->buf.obs; p.reward = h->buf.reward; p.terminated = h->buf.terminated;
    p.info_frame = h->buf.info_frame; p.info_misc = (uint32_t *)h->buf.info_misc;
    p.step_mask = h->buf.step_mask;
    p.tables = h->d_tables;
    p.first_env_index = h->cfg.first_env_index;
    p.n = h->cfg.num_envs; p.frame_skip = h->cfg.frame_skip; p.autoreset = h->cfg.autoreset;
    p.stale_intro = h->cfg.stale_intro_input;
    p.large_shape_min_envs = h->large_shape_min_envs;
    return p;
}

template <bool KF>
cudaError_t launch_step_k(const fg_config &c, int sm_count, cudaStream_t s, const Params &p) {
    const bool d = c.dense_reward != 0, m = p.step_mask != nullptr;
    if (c.p1_bot && c.p2_bot) return launch_step_d<KF, true, true>(d, m, sm_count, s, p);
    if (c.p1_bot) return launch_step_d<KF, true, false>(d, m, sm_count, s, p);
    if (c.p2_bot) return launch_step_d<KF, false, true>(d, m, sm_count, s, p);
    return launch_step_d<KF, false, false>(d, m, sm_count, s, p);
}

template <bool B1, bool B2>
cudaError_t launch_reset(int grid, cudaStream_t s, const Params &p) {
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(reset_kernel<B1, B2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Tables));
        if (e != cudaSuccess) return e;
        configured[dev & 63] = true;
    }
    reset_kernel<B1, B2><<<grid, kThreads, sizeof(Tables), s>>>(p);
    return cudaSuccess;
}

int check_bound(const fg_handle *h) {
    if (!h) return fail(FG_ERR_INVALID_ARGUMENT, "null handle%s");
    if (!h->bound) return fail(FG_ERR_NOT_BOUND, "fg_bind has not been called%s");
    return FG_OK;
}

}  // namespace

extern "C" {

int32_t fg_abi_version(void) { return FG_ABI_VERSION; }
const char *fg_last_error(void) { return g_err; }

int32_t fg_algorithmic_bytes_per_env_step(const fg_config *cfg) {
    if (!cfg) return 0;
    const bool rng = cfg->p1_bot || cfg->p2_bot;
    int state = 2 * 16 * (rng ? 4 : 3);                 // planes read + written
    int actions = (cfg->p1_bot ? 0 : 1) + (cfg->p2_bot ? 0 : 1);
    return state + actions + 32 /*obs*/ + 4 /*reward*/ + 1 /*terminated*/ + 4 /*info frame*/ + 4 /*info misc*/;
}

int32_t fg_create(const fg_config *cfg, fg_handle **out) {
    if (!cfg || !out) return fail(FG_ERR_INVALID_ARGUMENT, "null argument%s");
    if (cfg->struct_size != (int32_t)sizeof(fg_config)) return fail(FG_ERR_INVALID_ARGUMENT, "fg_config.struct_size mismatch%s");
    if (cfg->num_envs <= 0) return fail(FG_ERR_INVALID_ARGUMENT, "num_envs must be positive%s");
    if (cfg->frame_skip < 1 || cfg->frame_skip > 64) return fail(FG_ERR_INVALID_ARGUMENT, "frame_skip must be in [1, 64]%s");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(FG_ERR_NO_DEVICE, "no CUDA device: this library has no CPU fallback%s");
    }
    if (cfg->device < 0 || cfg->device >= ndev) return fail(FG_ERR_INVALID_ARGUMENT, "device ordinal out of range%s");
    CUDA_TRY(cudaSetDevice(cfg->device));
    fg_handle *h = new (std::nothrow) fg_handle();
    if (!h) return fail(FG_ERR_INVALID_STATE, "out of host memory%s");
    h->cfg = *cfg;
    h->bound = false;
    h->launches = 0;
    h->d_mask = nullptr;
    // developer / test knob: force the large CTA shapes onto small batches (or the small shape onto large ones)
    h->large_shape_min_envs = kLargeShapeMinEnvs;
    if (const char *v = getenv("FOOTSIES_B200_LARGE_SHAPE_MIN_ENVS")) h->large_shape_min_envs = atoi(v);
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, cfg->device));
    h->sm_count = prop.multiProcessorCount;
    Tables *host = new Tables();
    build_tables(*host);
    cudaError_t e = cudaMalloc(&h->d_tables, sizeof(Tables));
    if (e == cudaSuccess) e = cudaMemcpy(h->d_tables, host, sizeof(Tables), cudaMemcpyHostToDevice);
    delete host;
    if (e != cudaSuccess) { delete h; return fail(FG_ERR_CUDA, "table upload: %s", cudaGetErrorString(e)); }
    *out = h;
    return FG_OK;
}

void fg_destroy(fg_handle *h) {
    if (!h) return;
    cudaSetDevice(h->cfg.device);
    cudaFree(h->d_tables);
    if (h->d_mask) cudaFree(h->d_mask);
    delete h;
}

int32_t fg_bind(fg_handle *h, const fg_buffers *b) {
    if (!h || !b) return fail(FG_ERR_INVALID_ARGUMENT, "null argument%s");
    if (b->struct_size != (int32_t)sizeof(fg_buffers)) return fail(FG_ERR_INVALID_ARGUMENT, "fg_buffers.struct_size mismatch%s");
    for (int k = 0; k < FG_STATE_PLANES; k++)
        if (!b->state[k] || ((uintptr_t)b->state[k] & 15u)) return fail(FG_ERR_INVALID_ARGUMENT, "state planes must be non-null and 16-byte aligned%s");
    if (!b->stats || !b->obs || !b->reward || !b->terminated || !b->info_frame || !b->info_misc)
        return fail(FG_ERR_INVALID_ARGUMENT, "output buffers must be non-null%s");
    if (((uintptr_t)b->obs & 15u) || ((uintptr_t)b->info_misc & 3u) || ((uintptr_t)b->stats & 7u))
        return fail(FG_ERR_INVALID_ARGUMENT, "obs must be 16-byte, info_misc 4-byte, stats 8-byte aligned%s");
    if (!h->cfg.p1_bot && !b->actions_p1) return fail(FG_ERR_INVALID_ARGUMENT, "actions_p1 is required unless p1_bot%s");
    if (!h->cfg.p2_bot && !b->actions_p2) return fail(FG_ERR_INVALID_ARGUMENT, "actions_p2 is required unless p2_bot%s");
    h->buf = *b;
    h->bound = true;
    return FG_OK;
}

int32_t fg_seed(fg_handle *h, int64_t seed_base, const uint8_t *mask, void *stream) {
    if (int rc = check_bound(h)) return rc;
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    Params p = make_params(h);
    p.mask = mask; p.seed_base = seed_base;
    seed_kernel<<<grid_for(h, 8), kThreads, 0, (cudaStream_t)stream>>>(p);
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    return FG_OK;
}

int32_t fg_reset(fg_handle *h, const uint8_t *mask, void *stream) {
    if (int rc = check_bound(h)) return rc;
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    Params p = make_params(h);
    p.mask = mask;
    const int grid = grid_for(h, 4);
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t le;
    if (h->cfg.p1_bot && h->cfg.p2_bot) le = launch_reset<true, true>(grid, s, p);
    else if (h->cfg.p1_bot) le = launch_reset<true, false>(grid, s, p);
    else if (h->cfg.p2_bot) le = launch_reset<false, true>(grid, s, p);
    else le = launch_reset<false, false>(grid, s, p);
    CUDA_TRY(le);
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    return FG_OK;
}

int32_t fg_step(fg_handle *h, void *stream) {
    if (int rc = check_bound(h)) return rc;
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    const Params p = make_params(h);
    CUDA_TRY(h->cfg.frame_skip == 1 ? launch_step_k<false>(h->cfg, h->sm_count, (cudaStream_t)stream, p)
                                    : launch_step_k<true>(h->cfg, h->sm_count, (cudaStream_t)stream, p));
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    return FG_OK;
}

int32_t fg_step_host(fg_handle *h, const uint8_t *a1, const uint8_t *a2, float *obs, float *reward,
                     uint8_t *terminated, int32_t *info_frame, uint8_t *info_misc, void *stream) {
    if (int rc = check_bound(h)) return rc;
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    cudaStream_t s = (cudaStream_t)stream;
    const size_t n = (size_t)h->cfg.num_envs;
    if (!h->cfg.p1_bot) {
        if (!a1) return fail(FG_ERR_INVALID_ARGUMENT, "actions_p1 is required unless p1_bot%s");
        CUDA_TRY(cudaMemcpyAsync((void *)h->buf.actions_p1, a1, n, cudaMemcpyHostToDevice, s));
    }
    if (!h->cfg.p2_bot) {
        if (!a2) return fail(FG_ERR_INVALID_ARGUMENT, "actions_p2 is required unless p2_bot%s");
        CUDA_TRY(cudaMemcpyAsync((void *)h->buf.actions_p2, a2, n, cudaMemcpyHostToDevice, s));
    }
    if (int rc = fg_step(h, stream)) return rc;
    if (obs) CUDA_TRY(cudaMemcpyAsync(obs, h->buf.obs, n * 8 * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (reward) CUDA_TRY(cudaMemcpyAsync(reward, h->buf.reward, n * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (terminated) CUDA_TRY(cudaMemcpyAsync(terminated, h->buf.terminated, n, cudaMemcpyDeviceToHost, s));
    if (info_frame) CUDA_TRY(cudaMemcpyAsync(info_frame, h->buf.info_frame, n * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    if (info_misc) CUDA_TRY(cudaMemcpyAsync(info_misc, h->buf.info_misc, n * 4, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return FG_OK;
}

int32_t fg_reset_host(fg_handle *h, const uint8_t *mask, float *obs, int32_t *info_frame, uint8_t *info_misc, void *stream) {
    if (int rc = check_bound(h)) return rc;
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    cudaStream_t s = (cudaStream_t)stream;
    const size_t n = (size_t)h->cfg.num_envs;
    const uint8_t *dmask = nullptr;
    if (mask) {
        if (!h->d_mask) CUDA_TRY(cudaMalloc(&h->d_mask, n));
        CUDA_TRY(cudaMemcpyAsync(h->d_mask, mask, n, cudaMemcpyHostToDevice, s));
        dmask = h->d_mask;
    }
    if (int rc = fg_reset(h, dmask, stream)) return rc;
    if (obs) CUDA_TRY(cudaMemcpyAsync(obs, h->buf.obs, n * 8 * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (info_frame) CUDA_TRY(cudaMemcpyAsync(info_frame, h->buf.info_frame, n * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    if (info_misc) CUDA_TRY(cudaMemcpyAsync(info_misc, h->buf.info_misc, n * 4, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return FG_OK;
}

int32_t fg_get_state(fg_handle *h, int32_t first, int32_t count, fg_env_state *out) {
    if (int rc = check_bound(h)) return rc;
    if (!out || first < 0 || count < 0 || (int64_t)first + count > h->cfg.num_envs)
        return fail(FG_ERR_INVALID_ARGUMENT, "env range out of bounds%s");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    CUDA_TRY(cudaDeviceSynchronize());
    std::vector<FgVec4> pl[FG_STATE_PLANES];
    for (int k = 0; k < FG_STATE_PLANES; k++) {
        pl[k].resize((size_t)count);
        CUDA_TRY(cudaMemcpy(pl[k].data(), (const FgVec4 *)h->buf.state[k] + first, sizeof(FgVec4) * (size_t)count, cudaMemcpyDeviceToHost));
    }
    for (int i = 0; i < count; i++)
        fg_decode_env(pl[FG_PLANE_F1][i], pl[FG_PLANE_F2][i], pl[FG_PLANE_ENV][i], pl[FG_PLANE_RNG][i], &out[i]);
    return FG_OK;
}

int32_t fg_set_state(fg_handle *h, int32_t first, int32_t count, const fg_env_state *in) {
    if (int rc = check_bound(h)) return rc;
    if (!in || first < 0 || count < 0 || (int64_t)first + count > h->cfg.num_envs)
        return fail(FG_ERR_INVALID_ARGUMENT, "env range out of bounds%s");
    std::vector<FgVec4> pl[FG_STATE_PLANES];
    for (int k = 0; k < FG_STATE_PLANES; k++) pl[k].resize((size_t)count);
    for (int i = 0; i < count; i++)
        if (fg_encode_env(&in[i], &pl[FG_PLANE_F1][i], &pl[FG_PLANE_F2][i], &pl[FG_PLANE_ENV][i], &pl[FG_PLANE_RNG][i]))
            return fail(FG_ERR_INVALID_STATE, "state not representable (unknown action id, WIN/has_won, or field out of range)%s");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    CUDA_TRY(cudaDeviceSynchronize());
    for (int k = 0; k < FG_STATE_PLANES; k++)
        CUDA_TRY(cudaMemcpy((FgVec4 *)h->buf.state[k] + first, pl[k].data(), sizeof(FgVec4) * (size_t)count, cudaMemcpyHostToDevice));
    return FG_OK;
}

int32_t fg_read_stats(fg_handle *h, uint64_t *out, void *stream) {
    if (int rc = check_bound(h)) return rc;
    if (!out) return fail(FG_ERR_INVALID_ARGUMENT, "null output%s");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    CUDA_TRY(cudaMemcpyAsync(out, h->buf.stats, sizeof(uint64_t) * FG_STAT_COUNT, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    return FG_OK;
}

int64_t fg_launch_count(fg_handle *h) { return h ? h->launches : 0; }

}  // extern "C"
