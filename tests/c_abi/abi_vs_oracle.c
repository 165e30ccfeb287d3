/* abi_vs_oracle.c -- parity test of libfootsies_b200.so in plain C: no Python, no torch, nothing but the C ABI of
 * include/footsies_b200.h, the CUDA runtime for the caller's allocations, and the CPU oracle (oracle/footsies_oracle.h) as
 * the checker.  This is what a non-Python host of the reference (or a maintainer's C harness) sees of the library.
 *
 * TEST INFRASTRUCTURE (built and run by tests/test_c_abi.py): the oracle is only ever the checker.
 *
 *   abi_vs_oracle                run every case; exit 0 and print "ABI PARITY OK" when each field of each battle is
 *                                identical after every step (integers bit-exact, floats compared by their bit patterns)
 *   exit 3                       no CUDA device: the library has no CPU path and says so (fg_last_error)
 *
 * Cases: (1) device buffers the caller allocates, fg_bind / fg_seed / fg_reset / fg_step, random P1 vs BattleAI, ragged
 * batch; (2) self-play with fused frame-skip 3; (3) host buffers in, host buffers out: fg_step_host, and the packed
 * 16-byte records of fg_step_host_packed decoded here in C; (4) by_example (both bots) with a masked RESET + SEED. */
#include <cuda_runtime_api.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "footsies_b200.h"
#include "footsies_oracle.h"

#define CK(call)                                                                                       \
    do {                                                                                               \
        int32_t rc_ = (call);                                                                          \
        if (rc_ != FG_OK) { fprintf(stderr, "%s -> %d: %s\n", #call, rc_, fg_last_error()); exit(2); } \
    } while (0)
#define CU(call)                                                                                                \
    do {                                                                                                        \
        cudaError_t e_ = (call);                                                                                \
        if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e_)); exit(2); }         \
    } while (0)

static uint64_t lcg = 0x9e3779b97f4a7c15ull;
static uint32_t rnd(void) { lcg = lcg * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(lcg >> 33); }

typedef struct {
    int n;
    void *state[FG_STATE_PLANES];
    uint64_t *stats;
    uint8_t *a1, *a2, *terminated, *info_misc, *mask;
    float *obs, *reward;
    int32_t *info_frame;
    /* host mirrors */
    uint8_t *h_a1, *h_a2, *h_terminated, *h_info_misc, *h_mask;
    float *h_obs, *h_reward;
    int32_t *h_info_frame;
    fg_packed_result *h_packed;
    fo_trace *trace;
} bufs;

static void alloc_bufs(bufs *b, int n) {
    memset(b, 0, sizeof *b);
    b->n = n;
    for (int k = 0; k < FG_STATE_PLANES; k++) CU(cudaMalloc(&b->state[k], (size_t)n * FG_STATE_PLANE_BYTES_PER_ENV));
    CU(cudaMalloc((void **)&b->stats, FG_STAT_COUNT * sizeof(uint64_t)));
    CU(cudaMemset(b->stats, 0, FG_STAT_COUNT * sizeof(uint64_t)));
    CU(cudaMalloc((void **)&b->a1, n)); CU(cudaMalloc((void **)&b->a2, n)); CU(cudaMalloc((void **)&b->mask, n));
    CU(cudaMalloc((void **)&b->obs, (size_t)n * 32)); CU(cudaMalloc((void **)&b->reward, (size_t)n * 4));
    CU(cudaMalloc((void **)&b->terminated, n)); CU(cudaMalloc((void **)&b->info_frame, (size_t)n * 4));
    CU(cudaMalloc((void **)&b->info_misc, (size_t)n * 4));
    b->h_a1 = malloc(n); b->h_a2 = malloc(n); b->h_mask = malloc(n); b->h_terminated = malloc(n);
    b->h_info_misc = malloc((size_t)n * 4); b->h_obs = malloc((size_t)n * 32); b->h_reward = malloc((size_t)n * 4);
    b->h_info_frame = malloc((size_t)n * 4); b->h_packed = malloc((size_t)n * sizeof(fg_packed_result));
    b->trace = calloc(n, sizeof(fo_trace));
}
static void free_bufs(bufs *b) {
    for (int k = 0; k < FG_STATE_PLANES; k++) cudaFree(b->state[k]);
    cudaFree(b->stats); cudaFree(b->a1); cudaFree(b->a2); cudaFree(b->mask); cudaFree(b->obs); cudaFree(b->reward);
    cudaFree(b->terminated); cudaFree(b->info_frame); cudaFree(b->info_misc);
    free(b->h_a1); free(b->h_a2); free(b->h_mask); free(b->h_terminated); free(b->h_info_misc); free(b->h_obs);
    free(b->h_reward); free(b->h_info_frame); free(b->h_packed); free(b->trace);
}
static void bind(fg_handle *h, bufs *b) {
    fg_buffers fb;
    memset(&fb, 0, sizeof fb);
    fb.struct_size = (int32_t)sizeof fb;
    for (int k = 0; k < FG_STATE_PLANES; k++) fb.state[k] = b->state[k];
    fb.stats = b->stats; fb.actions_p1 = b->a1; fb.actions_p2 = b->a2; fb.obs = b->obs; fb.reward = b->reward;
    fb.terminated = b->terminated; fb.info_frame = b->info_frame; fb.info_misc = b->info_misc;
    CK(fg_bind(h, &fb));
}
static void fetch_outputs(bufs *b) {
    const size_t n = (size_t)b->n;
    CU(cudaMemcpy(b->h_obs, b->obs, n * 32, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(b->h_reward, b->reward, n * 4, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(b->h_terminated, b->terminated, n, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(b->h_info_frame, b->info_frame, n * 4, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(b->h_info_misc, b->info_misc, n * 4, cudaMemcpyDeviceToHost));
}

static long long mismatches = 0;
static void bad(const char *where, int step, int env, const char *field, double got, double exp) {
    if (mismatches++ < 10) fprintf(stderr, "MISMATCH [%s] step %d env %d %s: library %.9g oracle %.9g\n", where, step, env, field, got, exp);
}
/* step outputs (host arrays) against the oracle's trace */
static void check_outputs(const char *where, int step, const bufs *b, int with_reward) {
    for (int i = 0; i < b->n; i++) {
        const fo_trace *t = &b->trace[i];
        if (memcmp(&b->h_obs[8 * i], t->obs, 32) != 0)
            for (int k = 0; k < 8; k++)
                if (memcmp(&b->h_obs[8 * i + k], &t->obs[k], 4) != 0) bad(where, step, i, "obs", b->h_obs[8 * i + k], t->obs[k]);
        if (with_reward && memcmp(&b->h_reward[i], &t->reward, 4) != 0) bad(where, step, i, "reward", b->h_reward[i], t->reward);
        if (with_reward && b->h_terminated[i] != (uint8_t)t->terminated) bad(where, step, i, "terminated", b->h_terminated[i], t->terminated);
        if (b->h_info_frame[i] != t->info_frame) bad(where, step, i, "info_frame", b->h_info_frame[i], t->info_frame);
        for (int k = 0; k < 2; k++) {
            if (b->h_info_misc[4 * i + k] != (uint8_t)t->info_action[k]) bad(where, step, i, "info_action", b->h_info_misc[4 * i + k], t->info_action[k]);
            if (b->h_info_misc[4 * i + 2 + k] != (uint8_t)t->info_hitstun[k]) bad(where, step, i, "info_hitstun", b->h_info_misc[4 * i + 2 + k], t->info_hitstun[k]);
        }
    }
}
/* the library's battle state (fg_get_state) against the oracle's: every field both sides define identically */
static void check_state(const char *where, int step, fg_handle *h, const bufs *b, int with_rng) {
    fg_env_state *s = malloc((size_t)b->n * sizeof *s);
    CK(fg_get_state(h, 0, b->n, s));
    for (int i = 0; i < b->n; i++) {
        const fo_trace *t = &b->trace[i];
        for (int p = 0; p < 2; p++) {
            const fg_fighter_state *a = &s[i].f[p];
            const fo_fighter_state *o = &t->f[p];
            if (memcmp(&a->pos_x, &o->pos_x, 4) != 0) bad(where, step, i, "pos_x", a->pos_x, o->pos_x);
            if (memcmp(&a->velocity_x, &o->velocity_x, 4) != 0) bad(where, step, i, "velocity_x", a->velocity_x, o->velocity_x);
#define F(x) if (a->x != o->x) bad(where, step, i, #x, a->x, o->x)
            F(action_id); F(action_frame); F(hitstun); F(guard); F(vital); F(hit_count); F(buffer_id); F(reserve_id);
            F(is_input_backward); F(is_reserve_prox); F(shake); F(has_won); F(input0); F(attack_run);
#undef F
        }
        if (s[i].frame != t->frame) bad(where, step, i, "frame", s[i].frame, t->frame);
        if (s[i].done != t->terminated) bad(where, step, i, "done", s[i].done, t->terminated);
        for (int p = 0; p < 2; p++)
            if (s[i].recorded_input[p] != t->recorded_input[p]) bad(where, step, i, "recorded_input", s[i].recorded_input[p], t->recorded_input[p]);
        if (with_rng && memcmp(s[i].rng_state, t->rng_state, 16) != 0) bad(where, step, i, "rng_state", s[i].rng_state[3], t->rng_state[3]);
    }
    free(s);
}
static void check_stats(const char *where, fg_handle *h, fo_batch *o) {
    uint64_t st[FG_STAT_COUNT];
    int64_t os[FO_STAT_COUNT];
    double ret;
    CK(fg_read_stats(h, st, NULL));
    fo_stats(o, os, &ret);
    const int map[][2] = { { FG_STAT_EPISODES, FO_STAT_EPISODES }, { FG_STAT_P1_WINS, FO_STAT_P1_WINS }, { FG_STAT_P2_WINS, FO_STAT_P2_WINS },
                           { FG_STAT_DOUBLE_KO, FO_STAT_DOUBLE_KO }, { FG_STAT_EPISODE_FRAMES, FO_STAT_FRAMES },
                           { FG_STAT_P1_SPECIALS, FO_STAT_P1_SPECIALS }, { FG_STAT_P1_SPECIALS_NEUTRAL, FO_STAT_P1_SPECIALS_NEUTRAL },
                           { FG_STAT_GUARD_BREAKS, FO_STAT_GUARD_BREAKS }, { FG_STAT_HITS, FO_STAT_HITS }, { FG_STAT_BLOCKS, FO_STAT_BLOCKS } };
    for (unsigned k = 0; k < sizeof map / sizeof map[0]; k++)
        if ((int64_t)st[map[k][0]] != os[map[k][1]]) bad(where, -1, map[k][0], "statistic", (double)st[map[k][0]], (double)os[map[k][1]]);
    if ((int64_t)st[FG_STAT_ENV_FRAMES] != fo_frames_simulated(o)) bad(where, -1, FG_STAT_ENV_FRAMES, "env_frames", (double)st[FG_STAT_ENV_FRAMES], (double)fo_frames_simulated(o));
    printf("  [%s] episodes %llu, P1 wins %llu, hits %llu, blocks %llu, guard breaks %llu, env-frames %llu\n", where,
           (unsigned long long)st[FG_STAT_EPISODES], (unsigned long long)st[FG_STAT_P1_WINS], (unsigned long long)st[FG_STAT_HITS],
           (unsigned long long)st[FG_STAT_BLOCKS], (unsigned long long)st[FG_STAT_GUARD_BREAKS], (unsigned long long)st[FG_STAT_ENV_FRAMES]);
}
/* sticky random inputs: held for a while, so that dashes, charged specials, blocks and guard breaks all happen */
static void draw_actions(uint8_t *a, int n, int first_step) {
    for (int i = 0; i < n; i++)
        if (first_step || rnd() % 7 == 0) a[i] = (uint8_t)(rnd() & 7u);
}

static fg_config config(int n, int p1_bot, int p2_bot, int dense, int frame_skip, int64_t first) {
    fg_config c;
    memset(&c, 0, sizeof c);
    c.struct_size = (int32_t)sizeof c; c.num_envs = n; c.device = 0; c.p1_bot = p1_bot; c.p2_bot = p2_bot;
    c.dense_reward = dense; c.frame_skip = frame_skip; c.autoreset = 1; c.stale_intro_input = 1; c.first_env_index = first;
    return c;
}
static fo_batch *oracle(int n, int p1_bot, int p2_bot, int dense, int64_t first, int64_t seed) {
    fo_config oc = { p1_bot, p2_bot, dense, 0, 1, 1 };
    fo_batch *o = fo_create(n, &oc, first);
    fo_seed(o, seed, NULL);
    return o;
}

/* cases 1, 2, 4: device buffers, fg_step */
static void case_device(const char *name, int n, int steps, int p1_bot, int p2_bot, int dense, int k, int with_masked_reset) {
    const int64_t first = 4242, seed = -17;
    fg_config c = config(n, p1_bot, p2_bot, dense, k, first);
    fg_handle *h;
    bufs b;
    CK(fg_create(&c, &h));
    alloc_bufs(&b, n);
    bind(h, &b);
    fo_batch *o = oracle(n, p1_bot, p2_bot, dense, first, seed);
    CK(fg_seed(h, seed, NULL, NULL));
    CK(fg_reset(h, NULL, NULL));
    fo_reset(o, NULL, b.trace);
    fetch_outputs(&b);
    check_outputs(name, -1, &b, 0);
    check_state(name, -1, h, &b, 1);
    for (int t = 0; t < steps; t++) {
        if (with_masked_reset && t % 97 == 50) {                     /* RESET + SEED on a third of the battles, mid-round */
            for (int i = 0; i < n; i++) b.h_mask[i] = rnd() % 3 == 0;
            CU(cudaMemcpy(b.mask, b.h_mask, n, cudaMemcpyHostToDevice));
            CK(fg_seed(h, 1000 + t, b.mask, NULL));
            CK(fg_reset(h, b.mask, NULL));
            fo_seed(o, 1000 + t, b.h_mask);
            fo_reset(o, b.h_mask, b.trace);
            fetch_outputs(&b);
            check_outputs(name, t, &b, 0);
        }
        draw_actions(b.h_a1, n, t == 0);
        draw_actions(b.h_a2, n, t == 0);
        CU(cudaMemcpy(b.a1, b.h_a1, n, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(b.a2, b.h_a2, n, cudaMemcpyHostToDevice));
        CK(fg_step(h, NULL));
        fo_step(o, b.h_a1, p2_bot ? NULL : b.h_a2, k, b.trace, 4);
        fetch_outputs(&b);
        check_outputs(name, t, &b, 1);
        if (t % 25 == 24 || t == steps - 1) check_state(name, t, h, &b, 1);
    }
    if (!with_masked_reset) check_stats(name, h, o);
    printf("%s: %d battles x %d steps (frame_skip %d), %lld launches: %s\n", name, n, steps, k, (long long)fg_launch_count(h),
           mismatches ? "MISMATCH" : "identical");
    fo_destroy(o);
    fg_destroy(h);
    free_bufs(&b);
}

/* case 3: host buffers in, host buffers out */
static void case_host(const char *name, int n, int steps) {
    const int64_t first = 7, seed = 5;
    fg_config c = config(n, 0, 1, 1, 1, first);
    fg_handle *h;
    bufs b;
    CK(fg_create(&c, &h));
    alloc_bufs(&b, n);
    bind(h, &b);
    fo_batch *o = oracle(n, 0, 1, 1, first, seed);
    CK(fg_seed(h, seed, NULL, NULL));
    CK(fg_reset_host(h, NULL, b.h_obs, b.h_info_frame, b.h_info_misc, NULL));
    fo_reset(o, NULL, b.trace);
    check_outputs(name, -1, &b, 0);
    float table[FG_PACKED_REWARD_TABLE_SIZE];
    int32_t table_n = 0;
    CK(fg_packed_reward_table(h, table, &table_n));
    for (int t = 0; t < steps; t++) {
        draw_actions(b.h_a1, n, t == 0);
        if (t & 1) {
            CK(fg_step_host(h, b.h_a1, NULL, b.h_obs, b.h_reward, b.h_terminated, b.h_info_frame, b.h_info_misc, NULL));
        } else {
            /* one 16-byte record per battle; layout: include/footsies_b200.h fg_packed_result */
            CK(fg_step_host_packed(h, b.h_a1, NULL, b.h_packed, NULL));
            for (int i = 0; i < n; i++) {
                const fg_packed_result *r = &b.h_packed[i];
                const uint32_t w0 = r->w0, w1 = r->w1;
                float *ob = &b.h_obs[8 * i];
                ob[0] = (float)(w0 & 3u); ob[1] = (float)((w0 >> 2) & 3u);
                ob[2] = (float)((w0 >> 4) & 15u); ob[3] = (float)((w0 >> 8) & 15u);
                ob[4] = (float)((w0 >> 12) & 63u); ob[5] = (float)((w0 >> 18) & 63u);
                ob[6] = r->position[0]; ob[7] = r->position[1];
                b.h_terminated[i] = (uint8_t)((w0 >> 24) & 1u);
                b.h_info_misc[4 * i] = (uint8_t)((w0 >> 25) & 7u); b.h_info_misc[4 * i + 1] = (uint8_t)((w0 >> 28) & 7u);
                b.h_info_misc[4 * i + 2] = (uint8_t)(w1 & 31u); b.h_info_misc[4 * i + 3] = (uint8_t)((w1 >> 5) & 31u);
                const uint32_t ri = (w1 >> 10) & 127u;
                if ((int32_t)ri >= table_n) bad(name, t, i, "reward index", ri, table_n);
                b.h_reward[i] = table[ri < FG_PACKED_REWARD_TABLE_SIZE ? ri : 0];
                b.h_info_frame[i] = (int32_t)((w1 >> 17) & 0x7fffu) - 1;
            }
        }
        fo_step(o, b.h_a1, NULL, 1, b.trace, 4);
        check_outputs(name, t, &b, 1);
    }
    check_state(name, steps - 1, h, &b, 1);
    check_stats(name, h, o);
    printf("%s: %d battles x %d steps through host buffers (fg_step_host / fg_step_host_packed alternating): %s\n", name, n, steps,
           mismatches ? "MISMATCH" : "identical");
    fo_destroy(o);
    fg_destroy(h);
    free_bufs(&b);
}

int main(int argc, char **argv) {
    /* optional argument: divide every case's step count (runs under compute-sanitizer) */
    const int div = argc > 1 && atoi(argv[1]) > 0 ? atoi(argv[1]) : 1;
    printf("libfootsies_b200 ABI version %d (header %d)\n", fg_abi_version(), FG_ABI_VERSION);
    if (fg_abi_version() != FG_ABI_VERSION) { fprintf(stderr, "ABI version mismatch\n"); return 2; }
    {
        fg_config c = config(16, 0, 1, 1, 1, 0);
        fg_handle *h = NULL;
        const int32_t rc = fg_create(&c, &h);
        if (rc == FG_ERR_NO_DEVICE) { fprintf(stderr, "fg_create: %s\n", fg_last_error()); return 3; }
        if (rc != FG_OK) { fprintf(stderr, "fg_create -> %d: %s\n", rc, fg_last_error()); return 2; }
        /* error behaviour: stepping an unbound handle is refused, not undefined */
        if (fg_step(h, NULL) != FG_ERR_NOT_BOUND) { fprintf(stderr, "fg_step on an unbound handle was not refused\n"); return 2; }
        fg_destroy(h);
    }
    case_device("device buffers, random P1 vs BattleAI", 5000, 400 / div, 0, 1, 1, 1, 0);
    case_device("device buffers, self-play, fused frame-skip 3, sparse reward", 3001, 250 / div, 0, 0, 0, 3, 0);
    case_host("host buffers, random P1 vs BattleAI", 4099, 300 / div);
    case_device("device buffers, by_example (both bots), masked RESET + SEED", 1537, 300 / div, 1, 1, 1, 1, 1);
    if (mismatches) { fprintf(stderr, "%lld mismatches\n", mismatches); return 1; }
    printf("ABI PARITY OK\n");
    return 0;
}
