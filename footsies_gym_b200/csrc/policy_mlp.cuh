#ifndef FOOTSIES_B200_POLICY_MLP_CUH
#define FOOTSIES_B200_POLICY_MLP_CUH
// policy_mlp.cuh -- device code of the 8-H-H-8 tanh MLP policy (footsies_gym_b200.rollout.MLPPolicy), shared by the
// per-step policy kernel (policy_kernel.cu) and the whole-horizon rollout kernel (rollout_kernel.cu):
// scale -> Linear(8, H) -> tanh -> Linear(H, H) -> tanh -> Linear(H, 8) -> log-softmax -> categorical sample.
//
// Mapping: four threads share one env (each owns a quarter of the hidden units; partial sums travel by warp shuffles,
// the 4 threads of an env are adjacent lanes) and every thread carries kEnvsPerThread envs, so that each weight word
// fetched from shared memory is used for several envs (the inference is bound by those fetches).  Every multiply-add
// is an explicit fmaf: the result does not depend on the translation unit's -fmad setting (the simulator half of the
// rollout kernel must be compiled with -fmad=false), so both kernels produce bit-identical logits.
// Randomness: a counter-based hash of (seed, step counter, env index) -- reproducible, no state to carry.
#include <cuda_runtime.h>
#include <stdint.h>

namespace fgp {

constexpr int kMaxHidden = 128;
constexpr int kEnvsPerThread = 2;

struct PolicyWeights {
    const float *scale;        // [8]
    const float *w1, *b1;      // [H, 8], [H]      row-major [out][in] like torch.nn.Linear
    const float *w2, *b2;      // [H, H], [H]
    const float *w3, *b3;      // [8, H], [8]
};

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

// Shared-memory layout of one layer's weights for the "4 threads per env" mapping: thread `part` owns the output units
// part * Q .. part * Q + Q - 1 and needs, for every input k, its Q weights as contiguous 128-bit words:
//   ws[(k * 4 + part) * P + j] = W[part * Q + j][k]       (P = Q + 4: the 4 parts of a quarter-warp hit disjoint banks,
// e.g. Q = 16: 16-float segments at a 20-float pitch start at banks 0, 20, 8, 28)
template <int H>
struct PolicySmem {
    static_assert(H % 16 == 0 && H <= kMaxHidden, "hidden size");
    static constexpr int Q = H / 4, P = Q + 4;
    static constexpr int kW1 = 0, kW2 = kW1 + 8 * 4 * P, kW3 = kW2 + H * 4 * P, kB1 = kW3 + 8 * 4 * P, kB2 = kB1 + H,
                         kB3 = kB2 + H, kScale = kB3 + 8, kFloats = kScale + 8;
    static constexpr size_t kBytes = sizeof(float) * kFloats;
};

// Stage the weights into shared memory (all threads of the CTA; the caller synchronises afterwards).
template <int H>
__device__ __forceinline__ void policy_stage(float *sm, const PolicyWeights &p, int tid, int nthreads) {
    using L = PolicySmem<H>;
    constexpr int Q = L::Q, P = L::P;
    float *w1 = sm + L::kW1, *w2 = sm + L::kW2, *w3 = sm + L::kW3;
    for (int i = tid; i < H * 8; i += nthreads) {
        const int u = i / 8, k = i % 8;            // W1[u][k]
        w1[(k * 4 + u / Q) * P + u % Q] = p.w1[i];
        const int o = i / H, c = i % H;            // W3[o][c]
        w3[(o * 4 + c / Q) * P + c % Q] = p.w3[i];
    }
    for (int i = tid; i < H * H / 4; i += nthreads) {
        const float4 v = reinterpret_cast<const float4 *>(p.w2)[i];   // W2[u][k .. k + 3]
        const int u = (4 * i) / H, k = (4 * i) % H;
        float *dst = w2 + (u / Q) * P + u % Q;
        dst[(k + 0) * 4 * P] = v.x; dst[(k + 1) * 4 * P] = v.y; dst[(k + 2) * 4 * P] = v.z; dst[(k + 3) * 4 * P] = v.w;
    }
    for (int i = tid; i < H; i += nthreads) { sm[L::kB1 + i] = p.b1[i]; sm[L::kB2 + i] = p.b2[i]; }
    if (tid < 8) { sm[L::kB3 + tid] = p.b3[tid]; sm[L::kScale + tid] = p.scale[tid]; }
}

// Logits of this thread's E envs from their raw observation rows x[q][0..7] (scaled here).  Must be called by all 32
// lanes of a warp; `part` = lane & 3; the 4 lanes of a group hold the same x and end up with the same logits.
// ROLLED: the loop over the four source quarters of layer 2 stays a loop (the shuffle source lane and the weight
// address are runtime values), which cuts the unrolled code of that layer four-fold; same arithmetic order either way.
template <int H, int E, bool ROLLED = false>
__device__ __forceinline__ void policy_logits(const float *sm, int part, float (&x)[E][8], float (&lg)[E][8]) {
    using L = PolicySmem<H>;
    constexpr int Q = L::Q, P = L::P;
    const float *w1 = sm + L::kW1, *w2 = sm + L::kW2, *w3 = sm + L::kW3;
    const float *b1 = sm + L::kB1, *b2 = sm + L::kB2, *b3 = sm + L::kB3, *sc = sm + L::kScale;
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int q = 0; q < E; q++) {
#pragma unroll
        for (int k = 0; k < 8; k++) x[q][k] = x[q][k] * sc[k];
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
                h1[q][4 * j4] = fmaf(v.x, x[q][k], h1[q][4 * j4]); h1[q][4 * j4 + 1] = fmaf(v.y, x[q][k], h1[q][4 * j4 + 1]);
                h1[q][4 * j4 + 2] = fmaf(v.z, x[q][k], h1[q][4 * j4 + 2]); h1[q][4 * j4 + 3] = fmaf(v.w, x[q][k], h1[q][4 * j4 + 3]);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < Q; j++) {
#pragma unroll
        for (int q = 0; q < E; q++) h1[q][j] = fast_tanh(h1[q][j]);
    }
    // layer 2: each thread owns Q outputs and needs all H inputs: the other quarters of h1 come over warp shuffles
    float h2[E][Q];
#pragma unroll
    for (int j = 0; j < Q; j++) {
#pragma unroll
        for (int q = 0; q < E; q++) h2[q][j] = b2[part * Q + j];
    }
#pragma unroll(ROLLED ? 1 : 4)
    for (int src = 0; src < 4; src++) {
#pragma unroll
        for (int k = 0; k < Q; k++) {
            float hk[E];
#pragma unroll
            for (int q = 0; q < E; q++) hk[q] = __shfl_sync(0xffffffffu, h1[q][k], (lane & 28) | src, 32);   // h1[src * Q + k]
            const float4 *w = reinterpret_cast<const float4 *>(w2 + ((src * Q + k) * 4 + part) * P);
#pragma unroll
            for (int j4 = 0; j4 < Q / 4; j4++) {
                const float4 v = w[j4];
#pragma unroll
                for (int q = 0; q < E; q++) {
                    h2[q][4 * j4] = fmaf(v.x, hk[q], h2[q][4 * j4]); h2[q][4 * j4 + 1] = fmaf(v.y, hk[q], h2[q][4 * j4 + 1]);
                    h2[q][4 * j4 + 2] = fmaf(v.z, hk[q], h2[q][4 * j4 + 2]); h2[q][4 * j4 + 3] = fmaf(v.w, hk[q], h2[q][4 * j4 + 3]);
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
                s[q] = fmaf(v.x, h2[q][4 * j4], s[q]); s[q] = fmaf(v.y, h2[q][4 * j4 + 1], s[q]);
                s[q] = fmaf(v.z, h2[q][4 * j4 + 2], s[q]); s[q] = fmaf(v.w, h2[q][4 * j4 + 3], s[q]);
            }
        }
#pragma unroll
        for (int q = 0; q < E; q++) {
            s[q] += __shfl_xor_sync(0xffffffffu, s[q], 1, 32);
            s[q] += __shfl_xor_sync(0xffffffffu, s[q], 2, 32);
            lg[q][o] = s[q] + b3[o];
        }
    }
}

// log-softmax + inverse-CDF sample from 8 logits with the 32-bit random word `rnd`; returns the action index 0..7 (= the
// input bitmask Left 1 | Right 2 | Attack 4, wrappers/action_comb_disc.py:13-18) and its log-probability.
__device__ __forceinline__ int policy_sample(const float (&lg)[8], uint32_t rnd, float &logp) {
    float m = lg[0];
#pragma unroll
    for (int o = 1; o < 8; o++) m = fmaxf(m, lg[o]);
    float e[8], z = 0.0f;
#pragma unroll
    for (int o = 0; o < 8; o++) { e[o] = __expf(lg[o] - m); z += e[o]; }
    const float u01 = (float)(rnd >> 8) * (1.0f / 16777216.0f);   // [0, 1)
    const float target = u01 * z;
    int a = 7;
    float c = 0.0f, la = lg[7];
#pragma unroll
    for (int o = 0; o < 8; o++) { c += e[o]; if (a == 7 && target < c) { a = o; la = lg[o]; } }
    logp = (la - m) - __logf(z);
    return a;
}

}  // namespace fgp
#endif
