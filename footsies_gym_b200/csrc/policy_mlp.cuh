#ifndef FOOTSIES_B200_POLICY_MLP_CUH
#define FOOTSIES_B200_POLICY_MLP_CUH
// policy_mlp.cuh -- device code of the 8-H-H-8 tanh MLP policy (footsies_gym_b200.rollout.MLPPolicy), shared by the
// per-step policy kernel (policy_kernel.cu) and the whole-horizon rollout kernel (rollout_kernel.cu):
// scale -> Linear(8, H) -> tanh -> Linear(H, H) -> tanh -> Linear(H, 8) -> log-softmax -> categorical sample.
//
// Mapping (CTA = W warps, W = 4 or 8): warp w owns hidden units w * H/W .. (w + 1) * H/W - 1 for all of the CTA's battles and every
// lane carries E battles (lane, lane + 32, ...).  Weight fetches are warp-wide shared-memory broadcasts (one wavefront
// serves 32 x E battles), activations cross between the warps through shared memory ([unit][battle], conflict-free),
// and the layer-2 loop over the inputs stays a loop (small code).  History: the first version put the 4 threads of a
// battle into one quarter-warp and exchanged activations by shuffles -- a 128-bit weight fetch then delivers only 4
// distinct words per wavefront (measured: 98 wavefronts per env-step, shared-memory pipe 61 % busy) and its fully
// unrolled layer 2 thrashed the instruction cache at 4 battles per thread.
// Every multiply-add is an explicit fma (packed pairs, __ffma2_rn): the result does not depend on the translation
// unit's -fmad setting, and both kernels produce bit-identical logits.
// Randomness: a counter-based hash of (seed, step counter, GLOBAL env index = first_env_index + local index) --
// reproducible, no state to carry, and independent of how the battles are sharded over GPUs, like the simulator itself.
#include <cuda_runtime.h>
#include <stdint.h>

namespace fgp {

constexpr int kMaxHidden = 128;
constexpr int kPolicyWarps = 4;                    // W of both kernels (they must agree: the layer-3 partial sums are per warp);
                                                   // measured W = 8: 16 384 envs 6.7 vs 6.5 us per step, 1 Mi envs 352 vs 264

struct PolicyWeights {
    const float *scale;        // [8]
    const float *w1, *b1;      // [H, 8], [H]      row-major [out][in] like torch.nn.Linear
    const float *w2, *b2;      // [H, H], [H]
    const float *w3, *b3;      // [8, H], [8]
};

__device__ __forceinline__ float fast_tanh(float x) {
    // tanh(x) = 1 - 2 / (exp(2x) + 1) on the two MUFU approximations (relative error ~1e-6, far below what PPO's
    // ratios resolve): 5 instructions.  Same value as 1.0f - __fdividef(2.0f, __expf(2.0f * x) + 1.0f), whose
    // denormal fix-ups (12 instructions) can never fire here: the divisor is >= 1, and exp(2x) flushed to 0 or +inf
    // saturates the result to -1 / +1 either way.
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.8853900432586669921875f));   // 2 * log2(e), = 2 * the constant __expf uses
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
    return fmaf(-2.0f, r, 1.0f);
}

__device__ __forceinline__ uint32_t hash3(uint64_t seed, uint64_t counter, uint64_t idx) {
    uint64_t z = seed + 0x9e3779b97f4a7c15ull * (counter * 0x100000001b3ull + idx + 1ull);   // splitmix64 finaliser
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return (uint32_t)((z ^ (z >> 31)) >> 32);
}

// The observation as the other player sees it (for a policy that drives P2 the way it would drive P1): the per-player
// fields swap and the positions change sign.  The action such a policy picks is mirrored back with policy_mirror_action.
__device__ __forceinline__ void policy_mirror_obs(float (&x)[8]) {
    const float g = x[0], m = x[2], f = x[4], p = x[6];
    x[0] = x[1]; x[1] = g; x[2] = x[3]; x[3] = m; x[4] = x[5]; x[5] = f; x[6] = -x[7]; x[7] = -p;
}
__device__ __forceinline__ int policy_mirror_action(int a) { return (a & 4) | ((a & 1) << 1) | ((a >> 1) & 1); }   // Left <-> Right

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

// Shared memory of one CTA: transposed weights, biases, observation scale, the hidden activations of NE battles and
// the four partial logit sums per battle.
template <int H, int NE, int W>
struct PolicySmemBcast {
    static_assert(H % 16 == 0 && H <= kMaxHidden && (W == 4 || W == 8) && (H / W) % 4 == 0, "hidden size / warps");
    static constexpr int Q = H / W;                 // hidden units per warp
    static constexpr int kW1T = 0, kW2T = kW1T + 8 * H, kW3T = kW2T + H * H, kB1 = kW3T + 8 * H, kB2 = kB1 + H, kB3 = kB2 + H,
                         kScale = kB3 + 8, kHid = kScale + 8, kPart = kHid + H * NE, kFloats = kPart + W * 8 * NE;
    static constexpr size_t kBytes = sizeof(float) * kFloats;
    // the block splits into the weights [0, kWeightFloats) and the activations (hid, part) behind them, so that a second
    // policy can bring its own weights and share the activation buffers
    static constexpr int kWeightFloats = kHid, kActFloats = kFloats - kHid;
};

// w1t[k][u] = W1[u][k], w2t[k][u] = W2[u][k], w3t[c][o] = W3[o][c] (input-major: the outputs of one input are contiguous).
template <int H, int NE, int W>
__device__ __forceinline__ void policy_stage_bcast(float *sm, const PolicyWeights &p, int tid, int nthreads) {
    using L = PolicySmemBcast<H, NE, W>;
    for (int i = tid; i < H * 8; i += nthreads) {
        const int u = i / 8, k = i % 8;            // W1[u][k]
        sm[L::kW1T + k * H + u] = p.w1[i];
        sm[L::kW3T + (i % H) * 8 + i / H] = p.w3[i];   // W3[o][c], i = o * H + c
    }
    for (int i = tid; i < H * H / 4; i += nthreads) {
        const float4 v = reinterpret_cast<const float4 *>(p.w2)[i];   // W2[u][k .. k + 3]
        const int u = (4 * i) / H, k = (4 * i) % H;
        float *dst = sm + L::kW2T + u;
        dst[(k + 0) * H] = v.x; dst[(k + 1) * H] = v.y; dst[(k + 2) * H] = v.z; dst[(k + 3) * H] = v.w;
    }
    for (int i = tid; i < H; i += nthreads) { sm[L::kB1 + i] = p.b1[i]; sm[L::kB2 + i] = p.b2[i]; }
    if (tid < 8) { sm[L::kB3 + tid] = p.b3[tid]; sm[L::kScale + tid] = p.scale[tid]; }
}

// All 32 W threads of the CTA.  sm = the policy's weight block, act = the activation block (shared between policies);
// x[e] = raw observation row of battle lane + 32 e.  On return (after the function's last
// __syncthreads) the partial logits of every battle are in shared memory: policy_logits_of() assembles them.
// The multiply-adds are issued as packed pairs (__ffma2_rn, Blackwell's FFMA2: two independent correctly rounded fp32
// fmas per instruction, i.e. bit-identical to two fmaf) over adjacent output units.  Measured (tools/probes/
// ffma2_probe.cu): FFMA2 does not raise the FMA pipe's rate -- 2.0 cycles per warp instruction per scheduler against 1.1
// for FFMA, ~127 lane-fmas per cycle per SM either way -- it halves the ISSUE slots the fmas take, which is what this loop
// is short of (weight and activation fetches share them): 7.6 -> 6.9 us per rollout step.
template <int H, int E, int W>
__device__ __forceinline__ void policy_partials_bcast(const float *sm, float *act, int warp, int lane, float (&x)[E][8]) {
    constexpr int NE = 32 * E;
    using L = PolicySmemBcast<H, NE, W>;
    constexpr int Q = L::Q;
    const float *w1t = sm + L::kW1T + warp * Q, *w2t = sm + L::kW2T + warp * Q, *w3t = sm + L::kW3T + warp * Q * 8;
    const float *b1 = sm + L::kB1 + warp * Q, *b2 = sm + L::kB2 + warp * Q, *sc = sm + L::kScale;
    float *hid = act, *part = act + (L::kPart - L::kHid);
#pragma unroll
    for (int e = 0; e < E; e++) {
#pragma unroll
        for (int k = 0; k < 8; k++) x[e][k] = x[e][k] * sc[k];
    }
    float2 acc[E][Q / 2];                           // acc[e][j2] = units 2 * j2, 2 * j2 + 1 of this warp's quarter
    // layer 1
#pragma unroll
    for (int j2 = 0; j2 < Q / 2; j2++) {
#pragma unroll
        for (int e = 0; e < E; e++) acc[e][j2] = make_float2(b1[2 * j2], b1[2 * j2 + 1]);
    }
#pragma unroll
    for (int k = 0; k < 8; k++) {
#pragma unroll
        for (int j4 = 0; j4 < Q / 4; j4++) {
            const float4 v = *reinterpret_cast<const float4 *>(w1t + k * H + 4 * j4);
#pragma unroll
            for (int e = 0; e < E; e++) {
                const float2 xk = make_float2(x[e][k], x[e][k]);
                acc[e][2 * j4] = __ffma2_rn(make_float2(v.x, v.y), xk, acc[e][2 * j4]);
                acc[e][2 * j4 + 1] = __ffma2_rn(make_float2(v.z, v.w), xk, acc[e][2 * j4 + 1]);
            }
        }
    }
#pragma unroll
    for (int j2 = 0; j2 < Q / 2; j2++) {
#pragma unroll
        for (int e = 0; e < E; e++) {
            hid[(warp * Q + 2 * j2) * NE + lane + 32 * e] = fast_tanh(acc[e][j2].x);
            hid[(warp * Q + 2 * j2 + 1) * NE + lane + 32 * e] = fast_tanh(acc[e][j2].y);
        }
    }
    __syncthreads();
    // layer 2: all H inputs from shared memory, this warp's Q outputs
#pragma unroll
    for (int j2 = 0; j2 < Q / 2; j2++) {
#pragma unroll
        for (int e = 0; e < E; e++) acc[e][j2] = make_float2(b2[2 * j2], b2[2 * j2 + 1]);
    }
#pragma unroll 4                                    // measured: 2 / 4 / 8 / 16 within 2 % of each other
    for (int k = 0; k < H; k++) {
        float2 hk[E];
#pragma unroll
        for (int e = 0; e < E; e++) { const float h = hid[k * NE + lane + 32 * e]; hk[e] = make_float2(h, h); }
#pragma unroll
        for (int j4 = 0; j4 < Q / 4; j4++) {
            const float4 v = *reinterpret_cast<const float4 *>(w2t + k * H + 4 * j4);
#pragma unroll
            for (int e = 0; e < E; e++) {
                acc[e][2 * j4] = __ffma2_rn(make_float2(v.x, v.y), hk[e], acc[e][2 * j4]);
                acc[e][2 * j4 + 1] = __ffma2_rn(make_float2(v.z, v.w), hk[e], acc[e][2 * j4 + 1]);
            }
        }
    }
#pragma unroll
    for (int j2 = 0; j2 < Q / 2; j2++) {
#pragma unroll
        for (int e = 0; e < E; e++) acc[e][j2] = make_float2(fast_tanh(acc[e][j2].x), fast_tanh(acc[e][j2].y));
    }
    // layer 3: this warp's quarter of every logit; pairs run over adjacent outputs (w3t[j][o] = W3[o][j]), every logit
    // still accumulates its quarter in ascending j from 0
    float2 s[E][4];
#pragma unroll
    for (int e = 0; e < E; e++) {
#pragma unroll
        for (int o2 = 0; o2 < 4; o2++) s[e][o2] = make_float2(0.0f, 0.0f);
    }
#pragma unroll
    for (int j = 0; j < Q; j++) {
        const float4 va = *reinterpret_cast<const float4 *>(w3t + j * 8), vb = *reinterpret_cast<const float4 *>(w3t + j * 8 + 4);
#pragma unroll
        for (int e = 0; e < E; e++) {
            const float h = (j & 1) ? acc[e][j >> 1].y : acc[e][j >> 1].x;
            const float2 hj = make_float2(h, h);
            s[e][0] = __ffma2_rn(make_float2(va.x, va.y), hj, s[e][0]); s[e][1] = __ffma2_rn(make_float2(va.z, va.w), hj, s[e][1]);
            s[e][2] = __ffma2_rn(make_float2(vb.x, vb.y), hj, s[e][2]); s[e][3] = __ffma2_rn(make_float2(vb.z, vb.w), hj, s[e][3]);
        }
    }
#pragma unroll
    for (int o2 = 0; o2 < 4; o2++) {
#pragma unroll
        for (int e = 0; e < E; e++) {
            part[(warp * 8 + 2 * o2) * NE + lane + 32 * e] = s[e][o2].x;
            part[(warp * 8 + 2 * o2 + 1) * NE + lane + 32 * e] = s[e][o2].y;
        }
    }
    __syncthreads();
}

// The 8 logits of battle `local` (0 .. NE - 1) from the partial sums left by policy_partials_bcast: a fixed pairwise tree
// over the warps' partial sums, then the bias.
template <int H, int NE, int W>
__device__ __forceinline__ void policy_logits_of(const float *sm, const float *act, int local, float (&lg)[8]) {
    using L = PolicySmemBcast<H, NE, W>;
    const float *part = act + (L::kPart - L::kHid), *b3 = sm + L::kB3;
#pragma unroll
    for (int o = 0; o < 8; o++) {
        float s[W];
#pragma unroll
        for (int w = 0; w < W; w++) s[w] = part[(w * 8 + o) * NE + local];
#pragma unroll
        for (int span = 1; span < W; span *= 2) {
#pragma unroll
            for (int w = 0; w < W; w += 2 * span) s[w] = s[w] + s[w + span];
        }
        lg[o] = s[0] + b3[o];
    }
}

}  // namespace fgp
#endif
