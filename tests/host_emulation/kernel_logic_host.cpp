// kernel_logic_host.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Compiles the device frame logic (footsies_gym_b200/csrc/frame_logic.cuh: the very functions the sm_100a step
// kernel inlines) for the host, so that the parity suite can run them against the CPU oracle on a machine without a
// GPU.  This is NOT a CPU fallback of the product: it is built and loaded by tests/ only, lives outside the package,
// and libfootsies_b200.so contains no host path.  The per-env control flow below mirrors step_kernel / reset_kernel /
// seed_kernel in footsies_kernels.cu line by line.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../footsies_gym_b200/csrc/tables_host.h"

using namespace fg;

struct he_handle {
    int n, p1_bot, p2_bot, dense, frame_skip, autoreset, stale, skip_unactionable = 0;
    long long first_env_index;
    Tables T;
    std::vector<FgVec4> pl[4];
    std::vector<float> obs, reward;
    std::vector<uint8_t> terminated, info_misc;
    std::vector<int32_t> info_frame;
    unsigned long long stats[FG_STAT_COUNT];
};

static void load_env(const he_handle *h, int i, Env &e) {
    const FgVec4 a = h->pl[0][i], b = h->pl[1][i], c = h->pl[2][i], r = h->pl[3][i];
    e.pos1 = u2f(a.x); e.vel1 = u2f(a.y); e.pk1 = a.z; e.hist1 = a.w;
    e.pos2 = u2f(b.x); e.vel2 = u2f(b.y); e.pk2 = b.z; e.hist2 = b.w;
    e.frame = (int32_t)c.x; e.misc = c.y; e.bq2 = c.z; e.bq1 = c.w;
    e.r0 = r.x; e.r1 = r.y; e.r2 = r.z; e.r3 = r.w;
}
static void store_env(he_handle *h, int i, const Env &e) {
    h->pl[0][i] = FgVec4{ f2u(e.pos1), f2u(e.vel1), e.pk1, e.hist1 };
    h->pl[1][i] = FgVec4{ f2u(e.pos2), f2u(e.vel2), e.pk2, e.hist2 };
    h->pl[2][i] = FgVec4{ (uint32_t)e.frame, e.misc, e.bq2, e.bq1 };
    if (h->p1_bot || h->p2_bot) h->pl[3][i] = FgVec4{ e.r0, e.r1, e.r2, e.r3 };
}
static void write_outputs(he_handle *h, int i, const Env &e, float reward, bool terminated) {
    StepOutputs o;
    make_outputs(e, o);
    memcpy(&h->obs[8 * (size_t)i], o.obs, sizeof o.obs);
    h->reward[i] = reward;
    h->terminated[i] = terminated ? 1 : 0;
    h->info_frame[i] = e.frame;
    memcpy(&h->info_misc[4 * (size_t)i], &o.info_misc, 4);
}
static void flush(he_handle *h, StatAcc &acc) {
    static const int map[3][4] = { { FG_STAT_EPISODES, FG_STAT_P1_WINS, FG_STAT_P2_WINS, FG_STAT_DOUBLE_KO },
                                   { -1, FG_STAT_HITS, FG_STAT_BLOCKS, FG_STAT_GUARD_BREAKS },
                                   { FG_STAT_P1_SPECIALS, FG_STAT_P1_SPECIALS_NEUTRAL, FG_STAT_RESETS, FG_STAT_ENV_FRAMES } };
    const uint32_t w[3] = { acc.a, acc.r, acc.s };
    for (int k = 0; k < 3; k++)
        for (int b = 0; b < 4; b++)
            if (map[k][b] >= 0) h->stats[map[k][b]] += (w[k] >> (8 * b)) & 255u;
    h->stats[FG_STAT_EPISODE_FRAMES] += acc.ep_frames;
    acc = StatAcc{ 0u, 0u, 0u, 0u };
}

template <bool B1, bool B2>
static void reset_t(he_handle *h, const uint8_t *mask) {
    for (int i = 0; i < h->n; i++) {
        if (mask && !mask[i]) continue;
        Env e;
        load_env(h, i, e);
        reset_env<B1, B2>(h->T, e, h->stale != 0);
        store_env(h, i, e);
        write_outputs(h, i, e, 0.0f, false);
        h->stats[FG_STAT_RESETS]++;
    }
}

// FUSED mirrors the kernel's KFUSED template parameter (frame_skip > 1), which also selects the request-LUT variant
template <bool B1, bool B2, bool DENSE, bool FUSED>
static void step_t(he_handle *h, const uint8_t *a1, const uint8_t *a2, const uint8_t *step_mask) {
    for (int i = 0; i < h->n; i++) {
        if (step_mask && !step_mask[i]) continue;
        StatAcc acc = { 0u, 0u, 0u, 0u };
        Env e;
        load_env(h, i, e);
        bool run = false;
        uint32_t in1 = 0u, in2 = 0u;
        if ((e.misc >> FGM_DONE_SHIFT) & 1u) {
            if (h->autoreset) {
                reset_env<B1, B2>(h->T, e, h->stale != 0);
                store_env(h, i, e);
                write_outputs(h, i, e, 0.0f, false);
                acc.s += 0x10000u;
            } else {
                h->reward[i] = 0.0f;
            }
        } else {
            run = true;
            in1 = B1 ? (e.misc >> FGM_ACTOR1_SHIFT) & 7u : a1[i] & 7u;
            in2 = B2 ? (e.misc >> FGM_ACTOR2_SHIFT) & 7u : a2[i] & 7u;
        }
        double reward = 0.0;
        bool terminal = false;
        for (int kk = 0; kk < h->frame_skip; kk++) {
            if (run && !terminal) {
                simulate_frame<B1, B2, DENSE, FUSED>(h->T, e, in1, in2, reward, terminal, acc);
                acc.s += 1u << 24;
                if (B1) in1 = (e.misc >> FGM_ACTOR1_SHIFT) & 7u;
                if (B2) in2 = (e.misc >> FGM_ACTOR2_SHIFT) & 7u;
            }
        }
        if (h->skip_unactionable) {                                     // fused FootsiesFrameSkipped, as in step_kernel
            while (run && !terminal && obs_is_skippable(e)) {
                for (int kk = 0; kk < h->frame_skip; kk++) {
                    if (!terminal) {
                        simulate_frame<B1, B2, DENSE, FUSED>(h->T, e, 0u, in2, reward, terminal, acc);
                        acc.s += 1u << 24;
                        if (B2) in2 = (e.misc >> FGM_ACTOR2_SHIFT) & 7u;
                    }
                }
                flush(h, acc);
            }
        }
        if (run) {
            store_env(h, i, e);
            write_outputs(h, i, e, (float)reward, terminal);
        }
        flush(h, acc);
    }
}

extern "C" {

he_handle *he_create(int n, int p1_bot, int p2_bot, int dense, int frame_skip, int autoreset, int stale, long long first_env_index) {
    he_handle *h = new he_handle();
    h->n = n; h->p1_bot = p1_bot; h->p2_bot = p2_bot; h->dense = dense; h->frame_skip = frame_skip;
    h->autoreset = autoreset; h->stale = stale; h->first_env_index = first_env_index;
    build_tables(h->T);
    for (int k = 0; k < 4; k++) h->pl[k].assign((size_t)n, FgVec4{ 0, 0, 0, 0 });
    h->obs.assign((size_t)n * 8, 0.0f); h->reward.assign(n, 0.0f); h->terminated.assign(n, 0);
    h->info_misc.assign((size_t)n * 4, 0); h->info_frame.assign(n, 0);
    memset(h->stats, 0, sizeof h->stats);
    return h;
}
void he_destroy(he_handle *h) { delete h; }
void he_set_skip_unactionable(he_handle *h, int flag) { h->skip_unactionable = flag; }
void he_seed(he_handle *h, long long seed_base, const uint8_t *mask) {
    for (int i = 0; i < h->n; i++) {
        if (mask && !mask[i]) continue;
        uint32_t s0 = (uint32_t)(int32_t)(seed_base + h->first_env_index + i);
        uint32_t s1 = s0 * 1812433253u + 1u, s2 = s1 * 1812433253u + 1u, s3 = s2 * 1812433253u + 1u;
        h->pl[3][i] = FgVec4{ s0, s1, s2, s3 };
    }
}
void he_reset(he_handle *h, const uint8_t *mask) {
    if (h->p1_bot && h->p2_bot) reset_t<true, true>(h, mask);
    else if (h->p1_bot) reset_t<true, false>(h, mask);
    else if (h->p2_bot) reset_t<false, true>(h, mask);
    else reset_t<false, false>(h, mask);
}
void he_step(he_handle *h, const uint8_t *a1, const uint8_t *a2, const uint8_t *step_mask) {
#define GO(B1, B2) do { \
        if (h->frame_skip > 1 || h->skip_unactionable) { if (h->dense) step_t<B1, B2, true, true>(h, a1, a2, step_mask); else step_t<B1, B2, false, true>(h, a1, a2, step_mask); } \
        else { if (h->dense) step_t<B1, B2, true, false>(h, a1, a2, step_mask); else step_t<B1, B2, false, false>(h, a1, a2, step_mask); } \
    } while (0)
    if (h->p1_bot && h->p2_bot) GO(true, true);
    else if (h->p1_bot) GO(true, false);
    else if (h->p2_bot) GO(false, true);
    else GO(false, false);
#undef GO
}
int he_get_state(he_handle *h, int first, int count, fg_env_state *out) {
    for (int i = 0; i < count; i++)
        fg_decode_env(h->pl[0][first + i], h->pl[1][first + i], h->pl[2][first + i], h->pl[3][first + i], &out[i]);
    return 0;
}
int he_set_state(he_handle *h, int first, int count, const fg_env_state *in) {
    for (int i = 0; i < count; i++)
        if (fg_encode_env(&in[i], &h->pl[0][first + i], &h->pl[1][first + i], &h->pl[2][first + i], &h->pl[3][first + i])) return -1;
    return 0;
}
float *he_obs(he_handle *h) { return h->obs.data(); }
float *he_reward(he_handle *h) { return h->reward.data(); }
uint8_t *he_terminated(he_handle *h) { return h->terminated.data(); }
int32_t *he_info_frame(he_handle *h) { return h->info_frame.data(); }
uint8_t *he_info_misc(he_handle *h) { return h->info_misc.data(); }
unsigned long long *he_stats(he_handle *h) { return h->stats; }

}  // extern "C"
