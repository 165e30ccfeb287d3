#ifndef FOOTSIES_B200_STEP_KERNEL_CUH
#define FOOTSIES_B200_STEP_KERNEL_CUH
// step_kernel.cuh -- the sm_100a step / reset / seed kernels of the batched FOOTSIES simulator and their launchers.
//
// One CUDA thread owns one battle.  The whole reference frame update (csrc/frame_logic.cuh) runs in registers
// between one 64-byte state load and one 64-byte state store per env (four 16-byte SoA planes, fully coalesced
// 128-bit accesses).  Frame data is pre-expanded per (action, frame) (frame_tables.h) and staged once per CTA into
// shared memory; CTAs are persistent (grid-stride over chunks of 256 envs) and the state planes of the next chunk
// stream into shared memory with TMA bulk copies while the current chunk is simulated.
// No tensor cores: nothing here is a contraction.  Bounds: the ALU pipe first, then HBM bandwidth (K = 1).
// Results go back with per-thread 128-bit stores on purpose: staging a chunk's state and output planes in shared memory and
// writing them with TMA bulk stores (cp.async.bulk.global.shared::cta, the mirror of the load side) was built and measured
// in round 2 -- tools/probes/bulk_store_experiment.patch, profiles/r02v_bulk_store_ab.log -- and LOST: 143.5 vs 116.3 us per
// launch at 4 Mi battles, 44.8 vs 35.1 at 1 Mi, 9.8 vs 7.6 at 65 536 (the warps of a group get coupled through the staging
// block and its elected thread, where now they only meet at the load barrier).
#include "device_once.h"
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <new>
#include <vector>

#include "tables_host.h"

namespace fgk {

using namespace fg;

// CTA shape of the step kernel.  A CTA is GROUPS independent pipeline groups of GROUP threads: each group walks its
// own sequence of GROUP-env chunks with its own TMA stages and mbarriers; all groups share the one copy of the tables
// in shared memory.  Three shapes are compiled (measured on B200, tools/probes/run_sizes.sh):
//   ShapeSmall        256 threads = 1 group, 2 stages, 3 CTAs/SM (85 registers): many small CTAs, best below ~0.75 Mi envs
//                     (65 536 envs, K = 4: 10.1 us vs 15.2 us with a 1024-thread shape)
//   ShapeLargeSingle  768 threads = 3 groups x 256, 3 stages, 1 CTA/SM, K = 1: 85 registers per thread instead of 64 (the
//                     frame update wants ~85: no spills, more instruction-level parallelism) beats the extra 8 warps of
//                     a 1024-thread CTA while HBM is the co-limiter (4 Mi envs: 128 us vs 136 us)
//   ShapeLargeFused   1024 threads = 4 groups x 256, 3 stages, 1 CTA/SM, K > 1: purely ALU-bound, occupancy wins
//                     (1 Mi envs, K = 4: 97.8 us vs 103.3 us with 768 threads)
// All large shapes keep one copy of the tables per SM and prefetch two chunks ahead per group.
template <int THREADS_, int GROUP_, int STAGES_, int MIN_BLOCKS_>
struct StepShape {
    static constexpr int kThreads = THREADS_, kGroupThreads = GROUP_, kGroups = THREADS_ / GROUP_, kStages = STAGES_,
                         kMinBlocks = MIN_BLOCKS_;
    static_assert(THREADS_ % GROUP_ == 0 && GROUP_ % 32 == 0, "groups are whole warps");
};
#if defined(FG_THREADS)   // developer override: one shape for every batch size
#ifndef FG_GROUP
#define FG_GROUP FG_THREADS
#endif
#ifndef FG_STAGES
#define FG_STAGES 2
#endif
#ifndef FG_BLOCKS_PER_SM
#define FG_BLOCKS_PER_SM 4
#endif
using ShapeSmall = StepShape<FG_THREADS, FG_GROUP, FG_STAGES, FG_BLOCKS_PER_SM>;
using ShapeLargeSingle = ShapeSmall;
using ShapeLargeFused = ShapeSmall;
#else
using ShapeSmall = StepShape<256, 256, 2, 3>;
using ShapeLargeSingle = StepShape<768, 256, 3, 1>;
using ShapeLargeFused = StepShape<1024, 256, 3, 1>;
#endif
constexpr int kLargeShapeMinEnvs = 768 * 1024;
// developer switches for A/B measurements (tools/probes/build_variant.sh <tag> -DFG_...=0|1)
#ifndef FG_SKIP_RNG_STORE
#define FG_SKIP_RNG_STORE 1
#endif
#ifndef FG_PDL
#define FG_PDL 1
#endif
#ifndef FG_REVERSE
#define FG_REVERSE 1
#endif
constexpr int kThreads = 256;                    // reset / seed kernels
constexpr uint32_t kFull = 0xffffffffu;

struct Params {
    uint4 *pl_f1, *pl_f2, *pl_env, *pl_rng;
    unsigned long long *stats;
    const uint8_t *act1, *act2;
    float4 *obs;
    float *reward;
    uint8_t *terminated;
    int32_t *info_frame;
    uint32_t *info_misc;
    const Tables *tables;
    const uint8_t *mask;   // reset / seed kernels
    const uint8_t *step_mask;
    long long seed_base, first_env_index;
    int n, frame_skip, autoreset, stale_intro;
    int skip_unactionable;      // fused FootsiesFrameSkipped (KFUSED kernels only)
    int large_shape_min_envs;   // host side only: batch size from which the large CTA shapes are launched
    int reverse;                // walk the chunks from the last to the first (alternates per launch, see step_kernel)
    int pdl;                    // launched with programmatic stream serialization: dependent data only after griddepcontrol.wait
    int pdl_min_envs;           // host side only: batch size from which launches use it
};

__device__ __forceinline__ void write_outputs(const Params &p, int i, const Env &e, float reward, bool terminated) {
    StepOutputs o;
    make_outputs(e, o);
    p.obs[2 * (size_t)i] = make_float4(o.obs[0], o.obs[1], o.obs[2], o.obs[3]);
    p.obs[2 * (size_t)i + 1] = make_float4(o.obs[4], o.obs[5], o.obs[6], o.obs[7]);
    p.reward[i] = reward;
    p.terminated[i] = terminated ? 1 : 0;
    p.info_frame[i] = e.frame;
    p.info_misc[i] = o.info_misc;
}

__device__ __forceinline__ void load_tables(Tables *dst, const Tables *src) {
    const uint4 *s = reinterpret_cast<const uint4 *>(src);
    uint4 *d = reinterpret_cast<uint4 *>(dst);
    for (int k = threadIdx.x; k < (int)(sizeof(Tables) / 16); k += blockDim.x) d[k] = s[k];
}

template <bool WITH_RNG>
__device__ __forceinline__ void load_env(const Params &p, int i, Env &e) {
    const uint4 a = p.pl_f1[i], b = p.pl_f2[i], c = p.pl_env[i];
    e.pos1 = u2f(a.x); e.vel1 = u2f(a.y); e.pk1 = a.z; e.hist1 = a.w;
    e.pos2 = u2f(b.x); e.vel2 = u2f(b.y); e.pk2 = b.z; e.hist2 = b.w;
    e.frame = (int32_t)c.x; e.misc = c.y; e.bq2 = c.z; e.bq1 = c.w;
    if (WITH_RNG) { const uint4 r = p.pl_rng[i]; e.r0 = r.x; e.r1 = r.y; e.r2 = r.z; e.r3 = r.w; }
    e.drew = false;
}
template <bool WITH_RNG>
__device__ __forceinline__ void store_env(const Params &p, int i, const Env &e) {
    p.pl_f1[i] = make_uint4(f2u(e.pos1), f2u(e.vel1), e.pk1, e.hist1);
    p.pl_f2[i] = make_uint4(f2u(e.pos2), f2u(e.vel2), e.pk2, e.hist2);
    p.pl_env[i] = make_uint4((uint32_t)e.frame, e.misc, e.bq2, e.bq1);
    // the bot draws once every few dozen frames: an unchanged RNG plane entry is not written back (saves ~9 % of the
    // launch's DRAM traffic)
    if (WITH_RNG && (!FG_SKIP_RNG_STORE || e.drew)) p.pl_rng[i] = make_uint4(e.r0, e.r1, e.r2, e.r3);
}

// Warp-cooperative fold of the packed per-thread counters into the CTA's shared-memory vector.
__device__ __forceinline__ void flush_stats(StatAcc &acc, unsigned long long *s_stats, int lane) {
    const uint32_t w[3] = { acc.a, acc.r, acc.s };
    // (word, byte lane) -> statistic index; -1 = unused
    const int map[3][4] = { { FG_STAT_EPISODES, FG_STAT_P1_WINS, FG_STAT_P2_WINS, FG_STAT_DOUBLE_KO },
                            { -1, FG_STAT_HITS, FG_STAT_BLOCKS, FG_STAT_GUARD_BREAKS },
                                            { FG_STAT_P1_SPECIALS, FG_STAT_P1_SPECIALS_NEUTRAL, FG_STAT_RESETS, FG_STAT_ENV_FRAMES } };
#pragma unroll
    for (int k = 0; k < 3; k++)
#pragma unroll
        for (int b = 0; b < 4; b++) {
            if (map[k][b] < 0) continue;
            const uint32_t tot = __reduce_add_sync(kFull, (w[k] >> (8 * b)) & 255u);
            if (lane == 0 && tot) atomicAdd(&s_stats[map[k][b]], (unsigned long long)tot);
        }
    // episode lengths: two 16-bit halves so that the 32-lane sums cannot overflow
    const uint32_t lo = __reduce_add_sync(kFull, acc.ep_frames & 0xffffu), hi = __reduce_add_sync(kFull, acc.ep_frames >> 16);
    if (lane == 0 && (lo | hi)) atomicAdd(&s_stats[FG_STAT_EPISODE_FRAMES], ((unsigned long long)hi << 16) + lo);
    acc.a = 0u; acc.r = 0u; acc.s = 0u; acc.ep_frames = 0u;
}

// ---- TMA (1-D bulk async copy) + mbarrier helpers: the state planes of the NEXT chunk of 256 envs stream into
//      shared memory while the current chunk is being simulated ----
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <class SH, int PLANES>
struct __align__(128) StepSmem {
    Tables T;
    uint4 stage[SH::kGroups][SH::kStages][PLANES][SH::kGroupThreads];
    uint64_t full_bar[SH::kGroups][SH::kStages], empty_bar[SH::kGroups][SH::kStages];
    unsigned long long stats[FG_STAT_COUNT];
};

// FootsiesEnv.step for every env: up to K fused fight frames, or the reset of a finished env (autoreset).
// Persistent CTAs walk chunks of 256 consecutive envs; chunk c+grid is prefetched by one elected thread with
// 3-4 bulk copies of 4 KB (one per state plane) while chunk c is simulated out of registers.
template <class SH, bool KFUSED, bool P1BOT, bool P2BOT, bool DENSE, bool MASKED>
__global__ void __launch_bounds__(SH::kThreads, SH::kMinBlocks) step_kernel(const Params p) {
    constexpr bool kRng = P1BOT || P2BOT;
    constexpr int kPlanes = kRng ? 4 : 3;
    constexpr int kGroupThreads = SH::kGroupThreads, kGroups = SH::kGroups, kStages = SH::kStages;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    StepSmem<SH, kPlanes> &S = *reinterpret_cast<StepSmem<SH, kPlanes> *>(smem_raw);
    const Tables &T = S.T;
    const int lane = threadIdx.x & 31;
    const int g = threadIdx.x / kGroupThreads, lt = threadIdx.x % kGroupThreads;   // pipeline group, thread within it
    const int num_chunks = (p.n + kGroupThreads - 1) / kGroupThreads;
    const int full_chunks = p.n / kGroupThreads;                        // chunks that can be bulk-copied whole
    const int first = blockIdx.x * kGroups + g, stride = gridDim.x * kGroups;       // this group's chunk sequence
    const uint4 *const planes[4] = { p.pl_f1, p.pl_f2, p.pl_env, p.pl_rng };
    auto issue = [&](int chunk, int s) {                                // the group's elected thread only
        mbar_expect_tx(&S.full_bar[g][s], (uint32_t)(kPlanes * kGroupThreads * sizeof(uint4)));
#pragma unroll
        for (int k = 0; k < kPlanes; k++)
            tma_load_1d(S.stage[g][s][k], planes[k] + (size_t)chunk * kGroupThreads, (uint32_t)(kGroupThreads * sizeof(uint4)),
                        &S.full_bar[g][s]);
    };
    // Chunk order: launch k walks the chunks forwards, launch k + 1 backwards (p.reverse), so that a launch starts on the
    // battles whose state and output lines the previous launch left in the 126 MB L2.  Only the whole chunks swap places:
    // the ragged tail chunk (plain loads, no TMA stage, no barrier traffic) stays the LAST element of the sequence in both
    // directions, i.e. the final iteration of the group that owns it -- the stage / parity bookkeeping below relies on every
    // earlier iteration of a group being a staged one (found in round 2: a batch with a ragged tail AND several chunks per
    // group made the producer wait forever for a stage nobody had consumed, profiles/r02v_ragged_reverse_before_fix.log).
    auto chunk_of = [&](int c) { return (FG_REVERSE && p.reverse && c < full_chunks) ? full_chunks - 1 - c : c; };
    if (lt == 0) {
        for (int s = 0; s < kStages; s++) { mbar_init(&S.full_bar[g][s], 1u); mbar_init(&S.empty_bar[g][s], kGroupThreads / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // Programmatic dependent launch (p.pdl, large batches): this CTA may have become resident while the previous launch on
    // the stream was still draining.  What does not depend on it (barriers, the constant tables) is set up first; battle
    // state, actions and masks are only touched after griddepcontrol.wait, and the next launch is allowed to move in as soon
    // as SMs free up.  Without it (small batches: the kernel is a wave or less and its own start-up latency is what
    // counts) the first chunks are requested before the tables are staged, so that the two overlap.
    if (FG_PDL && p.pdl) {
        load_tables(&S.T, p.tables);
        if (threadIdx.x < FG_STAT_COUNT) S.stats[threadIdx.x] = 0ull;
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        asm volatile("griddepcontrol.wait;" ::: "memory");
    }
    if (lt == 0) {
        // prologue: the group's first kStages-1 chunks are in flight
        for (int j = 0; j < kStages - 1; j++) {
            const int cj = first + j * stride;
            if (cj < num_chunks && chunk_of(cj) < full_chunks) issue(chunk_of(cj), j);
        }
    }
    if (!(FG_PDL && p.pdl)) {
        load_tables(&S.T, p.tables);
        if (threadIdx.x < FG_STAT_COUNT) S.stats[threadIdx.x] = 0ull;
    }
    __syncthreads();
    StatAcc acc = { 0u, 0u, 0u, 0u };
    uint32_t frames_since_flush = 0u;
    // Actions are prefetched one chunk ahead.  (Measured alternatives, 4 Mi envs: riding the TMA barrier as a fifth
    // 256-byte bulk copy +3.5 %, register-less cp.async into per-thread slots +3 %, loading at the point of use +16 %.)
    uint32_t nin1 = 0u, nin2 = 0u;
    {
        const int i0 = (first < num_chunks ? chunk_of(first) : num_chunks) * kGroupThreads + lt;
        if (i0 < p.n) { if (!P1BOT) nin1 = p.act1[i0]; if (!P2BOT) nin2 = p.act2[i0]; }
    }
    int k = 0;
    for (int cl = first; cl < num_chunks; cl += stride, k++) {
        const int s = k % kStages;
        const int c = chunk_of(cl);
        const int i = c * kGroupThreads + lt;
        const bool staged = c < full_chunks;
        bool valid = staged || i < p.n;
        if (MASKED) valid = valid && p.step_mask[valid ? i : 0] != 0;
        if (lt == 0) {                                                  // producer: chunk k + kStages - 1 -> the stage read at k - 1
            const int cn = cl + (kStages - 1) * stride;
            if (cn < num_chunks && chunk_of(cn) < full_chunks) {
                const int sn = (k + kStages - 1) % kStages;
                if (k >= 1) mbar_wait(&S.empty_bar[g][sn], ((k - 1) / kStages) & 1);   // every warp has read chunk k-1
                issue(chunk_of(cn), sn);
            }
        }
        const uint32_t act1 = nin1, act2 = nin2;
        {
            const int cn1 = cl + stride;
            const int in = (cn1 < num_chunks ? chunk_of(cn1) : num_chunks) * kGroupThreads + lt;
            if (in < p.n) { if (!P1BOT) nin1 = p.act1[in]; if (!P2BOT) nin2 = p.act2[in]; }
        }
        Env e;
        bool run = false;
        uint32_t in1 = 0u, in2 = 0u;
        if (staged) {
            mbar_wait(&S.full_bar[g][s], (k / kStages) & 1);
            const uint4 a = S.stage[g][s][0][lt], b = S.stage[g][s][1][lt], cc = S.stage[g][s][2][lt];
            e.pos1 = u2f(a.x); e.vel1 = u2f(a.y); e.pk1 = a.z; e.hist1 = a.w;
            e.pos2 = u2f(b.x); e.vel2 = u2f(b.y); e.pk2 = b.z; e.hist2 = b.w;
            e.frame = (int32_t)cc.x; e.misc = cc.y; e.bq2 = cc.z; e.bq1 = cc.w;
            if (kRng) { const uint4 r = S.stage[g][s][kPlanes - 1][lt]; e.r0 = r.x; e.r1 = r.y; e.r2 = r.z; e.r3 = r.w; }
            e.drew = false;
            __syncwarp();
            if (lane == 0) mbar_arrive(&S.empty_bar[g][s]);
        } else if (valid) {
            load_env<kRng>(p, i, e);                                    // ragged tail chunk: plain loads
        }
        if (valid) {
            if ((e.misc >> FGM_DONE_SHIFT) & 1u) {
                if (p.autoreset) {                                      // next-step autoreset: this call only resets
                    reset_env<P1BOT, P2BOT>(T, e, p.stale_intro != 0);
                    store_env<kRng>(p, i, e);
                    write_outputs(p, i, e, 0.0f, false);
                    acc.s += 0x10000u;
                } else {
                    p.reward[i] = 0.0f;                                 // frozen until fg_reset
                }
            } else {
                run = true;
                in1 = P1BOT ? (e.misc >> FGM_ACTOR1_SHIFT) & 7u : act1 & 7u;
                in2 = P2BOT ? (e.misc >> FGM_ACTOR2_SHIFT) & 7u : act2 & 7u;
            }
        }
        double reward = 0.0;
        bool terminal = false;
        const int K = KFUSED ? p.frame_skip : 1;
        for (int kk = 0; kk < K; kk++) {
            if (run && !terminal) {
                simulate_frame<P1BOT, P2BOT, DENSE, KFUSED>(T, e, in1, in2, reward, terminal, acc);
                acc.s += 1u << 24;                                      // byte lane 3: env-frames simulated (<= 120 per flush)
                if (KFUSED) {
                    if (P1BOT) in1 = (e.misc >> FGM_ACTOR1_SHIFT) & 7u;
                    if (P2BOT) in2 = (e.misc >> FGM_ACTOR2_SHIFT) & 7u;
                }
            }
        }
        if (KFUSED && p.skip_unactionable) {
            // FootsiesFrameSkipped.step (wrappers/frame_skip.py:68-80): while the observation is one P1 cannot act on and
            // the battle is not over, take another env step (K frames) with P1's no-op input and add its reward.  The
            // loop is warp-synchronous (every lane goes round until no lane of the warp needs another step) so that the
            // statistics fold stays a warp-uniform decision.
            bool more = run && !terminal && obs_is_skippable(e);
            while (__any_sync(kFull, more)) {
                for (int kk = 0; kk < K; kk++) {
                    if (more && !terminal) {
                        simulate_frame<P1BOT, P2BOT, DENSE, KFUSED>(T, e, 0u, in2, reward, terminal, acc);
                        acc.s += 1u << 24;
                        if (P2BOT) in2 = (e.misc >> FGM_ACTOR2_SHIFT) & 7u;
                    }
                }
                more = more && !terminal && obs_is_skippable(e);
                frames_since_flush += (uint32_t)K;
                if (frames_since_flush >= kStatFlushFrames) { flush_stats(acc, S.stats, lane); frames_since_flush = 0u; }
            }
        }
        if (run) {
            store_env<kRng>(p, i, e);
            write_outputs(p, i, e, (float)reward, terminal);
        }
        frames_since_flush += (uint32_t)K;                              // uniform across the warp
        if (frames_since_flush >= kStatFlushFrames) { flush_stats(acc, S.stats, lane); frames_since_flush = 0u; }
    }
    flush_stats(acc, S.stats, lane);
    __syncthreads();
    if (threadIdx.x < FG_STAT_COUNT && S.stats[threadIdx.x]) atomicAdd(&p.stats[threadIdx.x], S.stats[threadIdx.x]);
}

// FootsiesEnv.reset / RESET command for the envs selected by mask (NULL = all).
template <bool P1BOT, bool P2BOT>
__global__ void __launch_bounds__(kThreads) reset_kernel(const Params p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Tables &T = *reinterpret_cast<Tables *>(smem_raw);
    load_tables(&T, p.tables);
    __syncthreads();
    constexpr bool kRng = P1BOT || P2BOT;
    unsigned long long resets = 0ull;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += gridDim.x * blockDim.x) {
        if (p.mask && !p.mask[i]) continue;
        Env e;
        load_env<kRng>(p, i, e);
        reset_env<P1BOT, P2BOT>(T, e, p.stale_intro != 0);
        store_env<kRng>(p, i, e);
        write_outputs(p, i, e, 0.0f, false);
        resets++;
    }
    if (resets) atomicAdd(&p.stats[FG_STAT_RESETS], resets);
}

// Random.InitState(seed_base + global env index) (BattleCore.cs:170-173)
static __global__ void __launch_bounds__(kThreads) seed_kernel(const Params p) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += gridDim.x * blockDim.x) {
        if (p.mask && !p.mask[i]) continue;
        uint32_t s0 = (uint32_t)(int32_t)(p.seed_base + p.first_env_index + i);
        uint32_t s1 = s0 * 1812433253u + 1u, s2 = s1 * 1812433253u + 1u, s3 = s2 * 1812433253u + 1u;
        p.pl_rng[i] = make_uint4(s0, s1, s2, s3);
    }
}

// The step kernel keeps its tables and the TMA stages in dynamic shared memory (> 48 KB): every instantiation is
// opted in once per process.
template <class SH, bool KF, bool B1, bool B2, bool D, bool M>
cudaError_t launch_step_shape(int sm_count, cudaStream_t s, const Params &p) {
    constexpr int kPlanes = (B1 || B2) ? 4 : 3;
    constexpr size_t bytes = sizeof(StepSmem<SH, kPlanes>);
    static DeviceOnceFlags configured;
    if (cudaError_t e = configure_once_per_device(configured, [] {
            return cudaFuncSetAttribute(step_kernel<SH, KF, B1, B2, D, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes); }))
        return e;
    const int want = (p.n + SH::kThreads - 1) / SH::kThreads, cap = sm_count * SH::kMinBlocks;
    const int grid = want < cap ? (want > 0 ? want : 1) : cap;
#if FG_PDL
    if (!p.pdl) {
        step_kernel<SH, KF, B1, B2, D, M><<<grid, SH::kThreads, bytes, s>>>(p);
        return cudaSuccess;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)SH::kThreads);
    cfg.dynamicSmemBytes = bytes;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, step_kernel<SH, KF, B1, B2, D, M>, p);
#else
    step_kernel<SH, KF, B1, B2, D, M><<<grid, SH::kThreads, bytes, s>>>(p);
    return cudaSuccess;
#endif
}
template <bool KF, bool B1, bool B2, bool D, bool M>
cudaError_t launch_step(int sm_count, cudaStream_t s, const Params &p) {
    if (p.n < p.large_shape_min_envs) return launch_step_shape<ShapeSmall, KF, B1, B2, D, M>(sm_count, s, p);
    if (KF) return launch_step_shape<ShapeLargeFused, KF, B1, B2, D, M>(sm_count, s, p);
    return launch_step_shape<ShapeLargeSingle, KF, B1, B2, D, M>(sm_count, s, p);
}
template <bool KF, bool B1, bool B2>
cudaError_t launch_step_d(bool dense, bool masked, int sm_count, cudaStream_t s, const Params &p) {
    if (dense) return masked ? launch_step<KF, B1, B2, true, true>(sm_count, s, p) : launch_step<KF, B1, B2, true, false>(sm_count, s, p);
    return masked ? launch_step<KF, B1, B2, false, true>(sm_count, s, p) : launch_step<KF, B1, B2, false, false>(sm_count, s, p);
}
// The 64 step-kernel variants are compiled in eight translation units (step_instances.cu, one per (KFUSED, P1BOT, P2BOT))
// so that the build parallelises; everybody else only sees these declarations.
#ifndef FG_STEP_INSTANCE
extern template cudaError_t launch_step_d<false, false, false>(bool, bool, int, cudaStream_t, const Params &);
extern template cudaError_t launch_step_d<false, false, true>(bool, bool, int, cudaStream_t, const Params &);
extern template cudaError_t launch_step_d<false, true, false>(bool, bool, int, cudaStream_t, const Params &);
extern template cudaError_t launch_step_d<false, true, true>(bool, bool, int, cudaStream_t, const Params &);
extern template cudaError_t launch_step_d<true, false, false>(bool, bool, int, cudaStream_t, const Params &);
extern template cudaError_t launch_step_d<true, false, true>(bool, bool, int, cudaStream_t, const Params &);
extern template cudaError_t launch_step_d<true, true, false>(bool, bool, int, cudaStream_t, const Params &);
extern template cudaError_t launch_step_d<true, true, true>(bool, bool, int, cudaStream_t, const Params &);
#endif

}  // namespace fgk
#endif
