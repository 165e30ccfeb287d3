// policy_kernel.cu -- fused inference of a small MLP policy over the step kernel's observation tensor:
// scale -> Linear(8, H) -> tanh -> Linear(H, H) -> tanh -> Linear(H, 8) -> log-softmax -> categorical sample, one launch.
//
// BASELINE.json configs[4] (PPO rollout, a torch MLP policy consuming obs in place, 16 384 envs per GPU) is bound by
// the dozen tiny torch kernels a policy step costs (~110 us per step, against ~6 us for the simulator step).  This
// kernel replaces them for the 8-H-H-8 tanh MLP of footsies_gym_b200.rollout.MLPPolicy: weights are staged once per CTA
// in shared memory (H = 64: 21 KB) and read as 128-bit broadcasts; four threads share one env (each owns a quarter of
// the hidden units, partial sums are exchanged with warp shuffles) so that 16 384 envs still fill the machine; the
// sampled action is written as the uint8 input bitmask the step kernel is bound to, next to its log-probability, and
// (optionally) the observation row is copied into the rollout buffer on the way.
// Randomness: a counter-based hash of (seed, call counter, env index) -- reproducible, no state to carry.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/footsies_b200.h"

namespace {

constexpr int kPolThreads = 128;       // 32 envs x 4 threads per CTA
constexpr int kMaxHidden = 128;
constexpr int kEnvsPerThread = 2;

__device__ __forceinline__ float fast_tanh(float x) {
    // tanh(x) = 1 - 2 / (exp(2x) + 1); __expf keeps the relative error ~1e-6, far below what PPO's ratios resolve
    const float e = __expf(2.0f * x);
    return 1.0f - __fdividef(2.0f, e + 1.0f);
}

__device__ __forceinline__ uint32_t hash3(uint64_t seed, uint64_t counter, uint32_t idx) {
    uint64_t z = seed + 0x9e3779b97f4a7c15ull * (counter * 0x100000001b3ull + idx + 1ull);   // splitmix64 finaliser
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return (uint32_t)((z ^ (z >> 31)) >> 32);
}

struct PolicyParams {
    const float *obs;          // [n, 8]
    const float *scale;        // [8]
    const float *w1, *b1;      // [H, 8], [H]
    const float *w2, *b2;      // [H, H], [H]
    const float *w3, *b3;      // [8, H], [8]
    uint8_t *actions;          // [n]
    float *logp;               // [n] or null
    float *obs_copy;           // [n, 8] or null: rollout-buffer slot for this step's observations
    unsigned long long seed, counter;
    const unsigned long long *counter_base;   // optional device word added to `counter` (CUDA-graph replays bump it)
    int n, hidden;
};

// Shared-memory layout of one layer's weights for the "4 threads per env" mapping: thread `part` owns the output units
// part * Q .. part * Q + Q - 1 and needs, for every input k, its Q weights as contiguous 128-bit words:
//   ws[(k * 4 + part) * kPad + j] = W[part * Q + j][k]       (kPad = Q rounded up so that the 4 parts of a quarter-warp
// hit disjoint banks: 16-float segments at a 20-float pitch start at banks 0, 20, 8, 28)
template <int Q> struct Pitch { static constexpr int v = Q + 4; };

template <int H>
__global__ void __launch_bounds__(kPolThreads) policy_mlp_sample_kernel(const PolicyParams p) {
    static_assert(H % 16 == 0 && H <= kMaxHidden, "hidden size");
    constexpr int Q = H / 4;                       // hidden units owned by one of the 4 threads of an env
    constexpr int P = Pitch<Q>::v;
    extern __shared__ __align__(16) float sm[];
    float *w1 = sm;                                // [8 inputs][4 parts][P]
    float *w2 = w1 + 8 * 4 * P;                    // [H inputs][4 parts][P]
    float *w3 = w2 + H * 4 * P;                    // [8 outputs][4 parts][P]: W3[o][part * Q + j]
    float *b1 = w3 + 8 * 4 * P, *b2 = b1 + H, *b3 = b2 + H, *sc = b3 + 8;
    for (int i = threadIdx.x; i < H * 8; i += kPolThreads) {
        const int u = i / 8, k = i % 8;            // W1[u][k]
        w1[(k * 4 + u / Q) * P + u % Q] = p.w1[i];
        const int o = i / H, c = i % H;            // W3[o][c]
        w3[(o * 4 + c / Q) * P + c % Q] = p.w3[i];
    }
    for (int i = threadIdx.x; i < H * H / 4; i += kPolThreads) {
        const float4 v = reinterpret_cast<const float4 *>(p.w2)[i];   // W2[u][k .. k + 3]
        const int u = (4 * i) / H, k = (4 * i) % H;
        float *dst = w2 + (u / Q) * P + u % Q;
        dst[(k + 0) * 4 * P] = v.x; dst[(k + 1) * 4 * P] = v.y; dst[(k + 2) * 4 * P] = v.z; dst[(k + 3) * 4 * P] = v.w;
    }
    for (int i = threadIdx.x; i < H; i += kPolThreads) { b1[i] = p.b1[i]; b2[i] = p.b2[i]; }
    if (threadIdx.x < 8) { b3[threadIdx.x] = p.b3[threadIdx.x]; sc[threadIdx.x] = p.scale[threadIdx.x]; }
    __syncthreads();
    const unsigned long long counter = p.counter + (p.counter_base ? *p.counter_base : 0ull);
    const int part = threadIdx.x & 3;              // which quarter of the hidden units
    // E envs per thread: every weight fetched from shared memory is used E times (the kernel is bound by those fetches)
    constexpr int E = kEnvsPerThread;
    constexpr int envs_per_block = (kPolThreads / 4) * E;
    for (int base = blockIdx.x * envs_per_block; base < p.n; base += gridDim.x * envs_per_block) {
        int env[E];
        bool valid[E];
        float x[E][8];
#pragma unroll
        for (int q = 0; q < E; q++) {
            env[q] = base + q * (kPolThreads / 4) + (threadIdx.x >> 2);
            valid[q] = env[q] < p.n;
            const float4 a = valid[q] ? reinterpret_cast<const float4 *>(p.obs)[2 * (size_t)env[q]] : make_float4(0, 0, 0, 0);
            const float4 b = valid[q] ? reinterpret_cast<const float4 *>(p.obs)[2 * (size_t)env[q] + 1] : make_float4(0, 0, 0, 0);
            if (valid[q] && p.obs_copy && part == 0) {
                reinterpret_cast<float4 *>(p.obs_copy)[2 * (size_t)env[q]] = a;
                reinterpret_cast<float4 *>(p.obs_copy)[2 * (size_t)env[q] + 1] = b;
            }
            x[q][0] = a.x * sc[0]; x[q][1] = a.y * sc[1]; x[q][2] = a.z * sc[2]; x[q][3] = a.w * sc[3];
            x[q][4] = b.x * sc[4]; x[q][5] = b.y * sc[5]; x[q][6] = b.z * sc[6]; x[q][7] = b.w * sc[7];
        }
        // layer 1: this thread's quarter of h1 = tanh(b1 + W1 x)
        float h1[E][Q];
#pragma unroll
        for (int j = 0; j < Q; j++) {
#pragma unroll
            for (int q = 0; q < E; q++) h1[q][j] = b1[part * Q + j];
        }
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const float4 *w = reinterpret_cast<const float4 *>(w1 + (k * 4 + part) * P);
#pragma unroll
            for (int j4 = 0; j4 < Q / 4; j4++) {
                const float4 v = w[j4];
#pragma unroll
                for (int q = 0; q < E; q++) {
                    h1[q][4 * j4] += v.x * x[q][k]; h1[q][4 * j4 + 1] += v.y * x[q][k];
                    h1[q][4 * j4 + 2] += v.z * x[q][k]; h1[q][4 * j4 + 3] += v.w * x[q][k];
                }
            }
        }
#pragma unroll
        for (int j = 0; j < Q; j++) {
#pragma unroll
            for (int q = 0; q < E; q++) h1[q][j] = fast_tanh(h1[q][j]);
        }
        // layer 2: each thread owns Q outputs and needs all H inputs: the other quarters of h1 come over warp
        // shuffles (the 4 threads of an env are adjacent lanes)
        float h2[E][Q];
#pragma unroll
        for (int j = 0; j < Q; j++) {
#pragma unroll
            for (int q = 0; q < E; q++) h2[q][j] = b2[part * Q + j];
        }
#pragma unroll
        for (int src = 0; src < 4; src++) {
#pragma unroll
            for (int k = 0; k < Q; k++) {
                float hk[E];
#pragma unroll
                for (int q = 0; q < E; q++) hk[q] = __shfl_sync(0xffffffffu, h1[q][k], (threadIdx.x & 28) | src, 32);   // h1[src * Q + k]
                const float4 *w = reinterpret_cast<const float4 *>(w2 + ((src * Q + k) * 4 + part) * P);
#pragma unroll
                for (int j4 = 0; j4 < Q / 4; j4++) {
                    const float4 v = w[j4];
#pragma unroll
                    for (int q = 0; q < E; q++) {
                        h2[q][4 * j4] += v.x * hk[q]; h2[q][4 * j4 + 1] += v.y * hk[q];
                        h2[q][4 * j4 + 2] += v.z * hk[q]; h2[q][4 * j4 + 3] += v.w * hk[q];
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < Q; j++) {
#pragma unroll
            for (int q = 0; q < E; q++) h2[q][j] = fast_tanh(h2[q][j]);
        }
        // layer 3: partial logits over this thread's quarter of h2, then a butterfly over the 4 lanes
        float lg[E][8];
#pragma unroll
        for (int o = 0; o < 8; o++) {
            const float4 *w = reinterpret_cast<const float4 *>(w3 + (o * 4 + part) * P);
            float s[E];
#pragma unroll
            for (int q = 0; q < E; q++) s[q] = 0.0f;
#pragma unroll
            for (int j4 = 0; j4 < Q / 4; j4++) {
                const float4 v = w[j4];
#pragma unroll
                for (int q = 0; q < E; q++) {
                    s[q] += v.x * h2[q][4 * j4]; s[q] += v.y * h2[q][4 * j4 + 1];
                    s[q] += v.z * h2[q][4 * j4 + 2]; s[q] += v.w * h2[q][4 * j4 + 3];
                }
            }
#pragma unroll
            for (int q = 0; q < E; q++) {
                s[q] += __shfl_xor_sync(0xffffffffu, s[q], 1, 32);
                s[q] += __shfl_xor_sync(0xffffffffu, s[q], 2, 32);
                lg[q][o] = s[q] + b3[o];
            }
        }
        // log-softmax + inverse-CDF sample (identical on the 4 lanes; lane 0 of the group writes)
#pragma unroll
        for (int q = 0; q < E; q++) {
            float m = lg[q][0];
#pragma unroll
            for (int o = 1; o < 8; o++) m = fmaxf(m, lg[q][o]);
            float e[8], z = 0.0f;
#pragma unroll
            for (int o = 0; o < 8; o++) { e[o] = __expf(lg[q][o] - m); z += e[o]; }
            const float u01 = (float)(hash3(p.seed, counter, (uint32_t)env[q]) >> 8) * (1.0f / 16777216.0f);   // [0, 1)
            const float target = u01 * z;
            int a = 7;
            float c = 0.0f, la = lg[q][7];
#pragma unroll
            for (int o = 0; o < 8; o++) { c += e[o]; if (a == 7 && target < c) { a = o; la = lg[q][o]; } }
            if (valid[q] && part == 0) {
                p.actions[env[q]] = (uint8_t)a;
                if (p.logp) p.logp[env[q]] = (la - m) - __logf(z);
            }
        }
    }
}

thread_local char g_perr[256] = "";

}  // namespace

extern "C" {

const char *fg_policy_last_error(void) { return g_perr; }

int32_t fg_policy_mlp_sample(const float *obs, const float *scale, const float *w1, const float *b1, const float *w2,
                             const float *b2, const float *w3, const float *b3, int32_t hidden, int32_t num_envs,
                             uint64_t seed, uint64_t counter, const uint64_t *counter_base, uint8_t *actions, float *logp,
                             float *obs_copy, void *stream) {
    if (!obs || !scale || !w1 || !b1 || !w2 || !b2 || !w3 || !b3 || !actions || num_envs <= 0) {
        snprintf(g_perr, sizeof g_perr, "fg_policy_mlp_sample: null argument or empty batch");
        return FG_ERR_INVALID_ARGUMENT;
    }
    if (hidden != 32 && hidden != 64 && hidden != 128) {
        snprintf(g_perr, sizeof g_perr, "fg_policy_mlp_sample: hidden size must be 32, 64 or 128");
        return FG_ERR_INVALID_ARGUMENT;
    }
    PolicyParams p = { obs, scale, w1, b1, w2, b2, w3, b3, actions, logp, obs_copy, seed, counter,
                       (const unsigned long long *)counter_base, num_envs, hidden };
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int envs_per_block = (kPolThreads / 4) * kEnvsPerThread;
    int grid = (num_envs + envs_per_block - 1) / envs_per_block;
    if (grid > sms * 4) grid = sms * 4;            // persistent: the weight staging is amortised over several env groups
    const size_t pitch = (size_t)hidden / 4 + 4;   // Pitch<Q>
    const size_t bytes = sizeof(float) * ((8 + (size_t)hidden + 8) * 4 * pitch + 2 * (size_t)hidden + 16);
    cudaError_t e = cudaSuccess;
    cudaStream_t s = (cudaStream_t)stream;
#define FG_POLICY_LAUNCH(HH) do { \
        static bool configured[64] = {}; \
        if (!configured[dev & 63]) { e = cudaFuncSetAttribute(policy_mlp_sample_kernel<HH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes); configured[dev & 63] = (e == cudaSuccess); } \
        if (e == cudaSuccess) policy_mlp_sample_kernel<HH><<<grid, kPolThreads, bytes, s>>>(p); } while (0)
    if (hidden == 32) FG_POLICY_LAUNCH(32);
    else if (hidden == 64) FG_POLICY_LAUNCH(64);
    else FG_POLICY_LAUNCH(128);
#undef FG_POLICY_LAUNCH
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(g_perr, sizeof g_perr, "fg_policy_mlp_sample: %s", cudaGetErrorString(e));
        return FG_ERR_CUDA;
    }
    return FG_OK;
}

}  // extern "C"
