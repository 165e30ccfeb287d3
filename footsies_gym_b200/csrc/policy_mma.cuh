#ifndef FOOTSIES_B200_POLICY_MMA_CUH
#define FOOTSIES_B200_POLICY_MMA_CUH
// policy_mma.cuh -- the 8-H-H-8 tanh MLP policy on the warp-level tensor-core path (mma.sync m16n8k8, TF32 operands,
// fp32 accumulate) with the 3 x TF32 operand split, for H <= 64 (H = 128 keeps the FFMA2 path of policy_mlp.cuh).
//
// Why (round 2, measured -- profiles/r02_mma_probe.log, profiles/r02d_rollout_battles_per_lane.log): the FFMA2 version
// of the rollout kernel is bound by ISSUE SLOTS and shared-memory weight fetches, not by the FMA pipe (one battle per
// lane instead of two: 10.9 us per step instead of 6.4).  An FFMA2 retires 64 multiply-adds per instruction, an m16n8k8
// MMA 1024; mma.sync sustains 512 TF32 MAC per cycle per SM on B200, i.e. 171 fp32-grade MAC per cycle per SM after the
// three-product split against ~127 for FFMA / FFMA2 -- and needs 16 x fewer instructions and one 128-bit weight fetch per
// 3072 multiply-adds.  Accuracy: a = a_hi + a_lo with a_hi = tf32(a), a_lo = tf32(a - a_hi) (likewise b); the product is
// a_lo*b_hi + a_hi*b_lo + a_hi*b_hi accumulated in fp32 (the dropped a_lo*b_lo term is ~2^-22 relative): the logits agree
// with torch's fp32 result to ~1e-6, the test bar is 2e-5 in log-probability (tests/test_rollout.py).
//
// Mapping: ONE WARP owns 32 battles for the whole forward pass -- two 16-row M tiles -- and nothing is shared between
// warps but the read-only weight fragments, so the rollout kernel needs no CTA barrier inside its horizon loop.  With
// g = lane / 4, t = lane % 4 the m16n8k8 fragments are (PTX ISA, "Matrix fragments for mma.m16n8k8"):
//   A (16 x 8, row): a0 = (g, t)  a1 = (g + 8, t)  a2 = (g, t + 4)  a3 = (g + 8, t + 4)
//   B (8 x 8, col):  b0 = (k = t, n = g)  b1 = (k = t + 4, n = g)
//   C (16 x 8):      c0 = (g, 2t)  c1 = (g, 2t + 1)  c2 = (g + 8, 2t)  c3 = (g + 8, 2t + 1)
// A layer's C tile for units 8j .. 8j + 7 becomes the next layer's A fragment of k-step j WITHOUT leaving the registers:
// the k index of an MMA is a summation index, so its order is free as long as A and B agree -- A column t is taken to be
// unit 8j + 2t and column t + 4 unit 8j + 2t + 1 (a0 = c0, a1 = c2, a2 = c1, a3 = c3), and the weight fragments are staged
// with their rows permuted the same way.
#include "policy_mlp.cuh"

namespace fgp {

constexpr int kMmaMaxHidden = 64;

// Shared memory of one weight set: per (layer, k-step, n-tile) one float4 per lane = {b0_hi, b1_hi, b0_lo, b1_lo}
// (one conflict-free LDS.128 per fragment), then the biases and the observation scale.
template <int H>
struct PolicyMmaSmem {
    static_assert(H % 8 == 0 && H <= kMmaMaxHidden, "hidden size of the MMA path");
    static constexpr int NT = H / 8;                              // n-tiles of a hidden layer = k-steps of the next one
    static constexpr int kW1 = 0, kW2 = kW1 + NT * 32, kW3 = kW2 + NT * NT * 32, kFrags = kW3 + NT * 32;   // in float4
    static constexpr int kB1 = kFrags * 4, kB2 = kB1 + H, kB3 = kB2 + H, kScale = kB3 + 8, kFloats = kScale + 8;   // in floats
    static constexpr size_t kBytes = sizeof(float) * kFloats;
};

__device__ __forceinline__ float tf32_hi(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
// Activation split on the hot path: hi = x truncated to TF32 (one LOP3), lo = x - hi, exact in fp32 and handed to the MMA
// as it is (the tensor core reads the upper 19 bits of a TF32 operand).  sm_100a has no native cvt.rna.tf32: it is emulated
// in ~5 integer instructions, which made the two conversions per activation a quarter of the policy's instructions
// (16 384 battles: 4.46 -> 4.00 us per step); the weights, split once per launch, keep the rounded conversion.
// |x - hi - tf32(lo)| <= 2^-21 |x|.
// (Tried and dropped, profiles/r02n_rollout_halves.log: computing layer 2 in two halves of its n-tiles to halve the live
// accumulators and fit a third CTA per SM -- 4.43 us per step at 16 384 battles, 200 instead of 184 us at 1 Mi.)
__device__ __forceinline__ void tf32_split(float x, uint32_t &hi, uint32_t &lo) {
    hi = __float_as_uint(x) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// c += a * b at fp32-grade accuracy: small terms first
__device__ __forceinline__ void mma_3xtf32(float (&c)[4], const uint32_t (&ahi)[4], const uint32_t (&alo)[4], const float4 &b) {
    mma_tf32(c, alo, __float_as_uint(b.x), __float_as_uint(b.y));
    mma_tf32(c, ahi, __float_as_uint(b.z), __float_as_uint(b.w));
    mma_tf32(c, ahi, __float_as_uint(b.x), __float_as_uint(b.y));
}

// Stage one weight set (torch.nn.Linear layout: W[out][in]) as pre-split, pre-permuted B fragments.
template <int H>
__device__ __forceinline__ void policy_mma_stage(float *sm, const PolicyWeights &p, int tid, int nthreads) {
    using L = PolicyMmaSmem<H>;
    constexpr int NT = L::NT;
    float4 *frag = reinterpret_cast<float4 *>(sm);
    auto put = [&](int idx, float b0, float b1) {
        const float h0 = tf32_hi(b0), h1 = tf32_hi(b1);
        frag[idx] = make_float4(h0, h1, tf32_hi(b0 - h0), tf32_hi(b1 - h1));
    };
    // layer 1: k = observation feature (natural order: t, t + 4), n = unit 8 nt + g
    for (int i = tid; i < NT * 32; i += nthreads) {
        const int nt = i >> 5, lane = i & 31, g = lane >> 2, t = lane & 3;
        put(L::kW1 + i, p.w1[(8 * nt + g) * 8 + t], p.w1[(8 * nt + g) * 8 + t + 4]);
    }
    // layer 2: k-step kt, column t <-> input unit 8 kt + 2t, column t + 4 <-> 8 kt + 2t + 1; n = output unit 8 nt + g
    for (int i = tid; i < NT * NT * 32; i += nthreads) {
        const int lane = i & 31, g = lane >> 2, t = lane & 3, nt = (i >> 5) % NT, kt = (i >> 5) / NT;
        put(L::kW2 + i, p.w2[(8 * nt + g) * H + 8 * kt + 2 * t], p.w2[(8 * nt + g) * H + 8 * kt + 2 * t + 1]);
    }
    // layer 3: one n-tile (the 8 logits), n = g
    for (int i = tid; i < NT * 32; i += nthreads) {
        const int kt = i >> 5, lane = i & 31, g = lane >> 2, t = lane & 3;
        put(L::kW3 + i, p.w3[g * H + 8 * kt + 2 * t], p.w3[g * H + 8 * kt + 2 * t + 1]);
    }
    for (int i = tid; i < H; i += nthreads) { sm[L::kB1 + i] = p.b1[i]; sm[L::kB2 + i] = p.b2[i]; }
    if (tid < 8) { sm[L::kB3 + tid] = p.b3[tid]; sm[L::kScale + tid] = p.scale[tid]; }
}

// The 8 logits of the warp's 16 MT battles.  obs_rows: the battles' raw observation rows ([16 MT][8] floats, shared or
// global memory); lg_rows: [16 MT][8] floats of shared memory private to the warp, where row r receives battle r's logits
// (the caller reads its own row after the __syncwarp at the end).  MIRROR shows the network the observation as the other
// player sees it (policy_mirror_obs: per-player fields swap, positions change sign).
// MT = M tiles per warp: 2 (32 battles, every lane simulates one) or 1 (16 battles: twice the warps per SM for small batches).
template <int H, int MT, bool MIRROR>
__device__ __forceinline__ void policy_mma_logits(const float *sm, const float *obs_rows, float *lg_rows, int lane) {
    static_assert(MT == 1 || MT == 2, "one or two M tiles per warp");
    using L = PolicyMmaSmem<H>;
    constexpr int NT = L::NT;
    const float4 *frag = reinterpret_cast<const float4 *>(sm);
    const float *b1 = sm + L::kB1, *b2 = sm + L::kB2, *b3 = sm + L::kB3, *sc = sm + L::kScale;
    const int g = lane >> 2, t = lane & 3;
    // ---- A fragments of layer 1 from the observation rows (features t and t + 4 of rows g, g + 8 of each M tile) ----
    uint32_t xhi[MT][4], xlo[MT][4];
    {
        // mirrored observation: feature f comes from f ^ 1, positions (6, 7) change sign; the scale belongs to the feature
        // as the network sees it
        const int f0 = MIRROR ? (t ^ 1) : t, f1 = MIRROR ? ((t + 4) ^ 1) : t + 4;
        const float s0 = sc[t], s1 = (MIRROR && t >= 2) ? -sc[t + 4] : sc[t + 4];
#pragma unroll
        for (int m = 0; m < MT; m++) {
            const float *r0 = obs_rows + (16 * m + g) * 8, *r1 = r0 + 64;
            tf32_split(r0[f0] * s0, xhi[m][0], xlo[m][0]);
            tf32_split(r1[f0] * s0, xhi[m][1], xlo[m][1]);
            tf32_split(r0[f1] * s1, xhi[m][2], xlo[m][2]);
            tf32_split(r1[f1] * s1, xhi[m][3], xlo[m][3]);
        }
    }
    // ---- layer 1: h[m][nt] = tanh(x W1^T + b1), kept as C tiles ----
    float h[MT][NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; nt++) {
        const float4 b = frag[L::kW1 + nt * 32 + lane];
        const float2 bias = *reinterpret_cast<const float2 *>(b1 + 8 * nt + 2 * t);
#pragma unroll
        for (int m = 0; m < MT; m++) {
            h[m][nt][0] = bias.x; h[m][nt][1] = bias.y; h[m][nt][2] = bias.x; h[m][nt][3] = bias.y;
            mma_3xtf32(h[m][nt], xhi[m], xlo[m], b);
#pragma unroll
            for (int k = 0; k < 4; k++) h[m][nt][k] = fast_tanh(h[m][nt][k]);
        }
    }
    // ---- layer 2 ----
    float acc[MT][NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; nt++) {
        const float2 bias = *reinterpret_cast<const float2 *>(b2 + 8 * nt + 2 * t);
#pragma unroll
        for (int m = 0; m < MT; m++) { acc[m][nt][0] = bias.x; acc[m][nt][1] = bias.y; acc[m][nt][2] = bias.x; acc[m][nt][3] = bias.y; }
    }
#pragma unroll
    for (int kt = 0; kt < NT; kt++) {
        uint32_t ahi[MT][4], alo[MT][4];
#pragma unroll
        for (int m = 0; m < MT; m++) {                 // C tile kt -> A fragment of k-step kt: a0 = c0, a1 = c2, a2 = c1, a3 = c3
            tf32_split(h[m][kt][0], ahi[m][0], alo[m][0]);
            tf32_split(h[m][kt][2], ahi[m][1], alo[m][1]);
            tf32_split(h[m][kt][1], ahi[m][2], alo[m][2]);
            tf32_split(h[m][kt][3], ahi[m][3], alo[m][3]);
        }
#pragma unroll
        for (int nt = 0; nt < NT; nt++) {
            const float4 b = frag[L::kW2 + (kt * NT + nt) * 32 + lane];
#pragma unroll
            for (int m = 0; m < MT; m++) mma_3xtf32(acc[m][nt], ahi[m], alo[m], b);
        }
    }
    // ---- layer 3 on tanh(layer 2) ----
    float lg[MT][4];
    {
        const float2 bias = *reinterpret_cast<const float2 *>(b3 + 2 * t);
#pragma unroll
        for (int m = 0; m < MT; m++) { lg[m][0] = bias.x; lg[m][1] = bias.y; lg[m][2] = bias.x; lg[m][3] = bias.y; }
    }
#pragma unroll
    for (int kt = 0; kt < NT; kt++) {
        const float4 b = frag[L::kW3 + kt * 32 + lane];
#pragma unroll
        for (int m = 0; m < MT; m++) {
            uint32_t ahi[4], alo[4];
            tf32_split(fast_tanh(acc[m][kt][0]), ahi[0], alo[0]);
            tf32_split(fast_tanh(acc[m][kt][2]), ahi[1], alo[1]);
            tf32_split(fast_tanh(acc[m][kt][1]), ahi[2], alo[2]);
            tf32_split(fast_tanh(acc[m][kt][3]), ahi[3], alo[3]);
            mma_3xtf32(lg[m], ahi, alo, b);
        }
    }
    // ---- logits of row r to lg_rows[r][0 .. 7]: this lane holds columns 2t, 2t + 1 of rows 16 m + g and 16 m + g + 8 ----
    __syncwarp();                                     // the previous step's readers are done with lg_rows
#pragma unroll
    for (int m = 0; m < MT; m++) {
        *reinterpret_cast<float2 *>(lg_rows + (16 * m + g) * 8 + 2 * t) = make_float2(lg[m][0], lg[m][1]);
        *reinterpret_cast<float2 *>(lg_rows + (16 * m + g + 8) * 8 + 2 * t) = make_float2(lg[m][2], lg[m][3]);
    }
    __syncwarp();
}

}  // namespace fgp
#endif
