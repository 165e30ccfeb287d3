// state_codec.h -- compact per-env battle state (64 B/env in four 16-byte SoA planes) and its
// expanded form.  Shared by the CUDA kernels and the host-side fg_get_state / fg_set_state.
//
// Why this is enough (vs. the reference's 3 x 180 ints of input history per fighter, Fighter.cs:98-101):
//   * dash detection reads Left/Right of input[0..16] only (Fighter.cs:585-635 with dashAllowFrame 9: i <= 8,
//     j <= i + 8), and all it can learn from them is: how long ago the most recent direction-held frame was, which
//     directions it held and how long the run of direction-held frames ending there is -> a 217-state automaton
//     (tools/gen_kernel_tables.py build_dash_fsm), stored as its 8-bit state id
//   * the hold-release special reads "Attack held on input[1..59]" (Fighter.cs:569-583) -> a run length saturating at 59
//   * inputDown / inputUp are functions of input[0], input[1] (Fighter.cs:184-185).
// fg_get_state expands the automaton state into an equivalent Left/Right bit history (the run of `lastdir` frames
// ending `since` frames ago); fg_set_state reduces any bit history to its automaton state.
#ifndef FOOTSIES_B200_STATE_CODEC_H
#define FOOTSIES_B200_STATE_CODEC_H

#include <stdint.h>
#include "../../include/footsies_b200.h"
#include "frame_tables.h"

#if defined(__CUDACC__)
#define FG_HD __host__ __device__ __forceinline__
#else
#define FG_HD static inline
#endif

// plane indices
#define FG_PLANE_F1 0   // {pos_x bits, velocity_x bits, packed, input word}
#define FG_PLANE_F2 1
#define FG_PLANE_ENV 2  // {frame, misc, bot queue P2, bot queue P1}
#define FG_PLANE_RNG 3  // xorshift128 state

// packed fighter word.  The low 11 bits (frame | action << 6) are the index of the fighter's row in the
// [action][64 frames] frame table, so the table lookup needs no arithmetic.
#define FGP_FRAME_SHIFT 0    // 6 bits: currentActionFrame (every reachable frame is < 63; DEAD saturates)
#define FGP_ACT_SHIFT 6      // 5 bits: action index (moves.py order)
#define FGP_STUN_SHIFT 11    // 5 bits
#define FGP_GUARD_SHIFT 16   // 2 bits
#define FGP_VITAL_SHIFT 18   // 1 bit
#define FGP_HITCNT_SHIFT 19  // 1 bit (numberOfHit == 1 for every attack)
#define FGP_BUF_SHIFT 20     // 1 bit: bufferActionID == 110
#define FGP_RSV_SHIFT 21     // 1 bit: reserveDamageActionID == 310
#define FGP_INBACK_SHIFT 22  // 1 bit: isInputBackward
#define FGP_RPROX_SHIFT 23   // 1 bit: isReserveProximityGuard
#define FGP_SHAKE_SHIFT 24   // 4 bits sign-magnitude: |spriteShakePosition| in [24:27), negative in [27]
#define FGP_SHAKE_SIGN_SHIFT 27
// carry bits: facts about (action, frame) that the NEXT frame's request logic needs, copied from the row (row.w)
#define FGP_CARRY_END (1u << 28)     // the action is over as soon as the frame counter increments (Fighter.cs:90);
                                     // never set while in hit stun (the counter is frozen then)
#define FGP_CARRY_ALWAYS (1u << 29)  // alwaysCancelable
#define FGP_CARRY_NORMAL (1u << 30)  // N_ATTACK / B_ATTACK
#define FGP_CARRY_MASK (7u << 28)
#define FGP_ROW_MASK 0x7ffu
#define FGP_MAX_FRAME 63

// input word of a fighter (4th word of its plane)
#define FGH_DASH_SHIFT 0     // 8 bits: dash-detection automaton state (0 = COLD / cleared history)
#define FGH_ARUN_SHIFT 8     // 6 bits: attack run length (saturates at 59)

// misc env word (bits [9:12) unused)
#define FGM_P1MEM_SHIFT 0    // 8 bits, by_example only: the FightState P1's never-Reset() BattleAI recorded last, kept over a
                             // terminal frame for its first query of the next round: distance bucket [0:3) | opponent action [3:8)
#define FGM_P1CALLED_SHIFT 8 // 1 bit, by_example only: P1's BattleAI has been queried before (its fightStates are not null)
#define FGM_REC1_SHIFT 12    // 3 bits: last recorded input P1 (BattleCore.cs:593-607 stops recording after 18000 frames)
#define FGM_REC2_SHIFT 15
#define FGM_DONE_SHIFT 18    // battle over, waiting for reset
#define FGM_CUM_SHIFT 19     // 4 bits: dense-reward automaton index
#define FGM_ACTOR1_SHIFT 23  // 3 bits: input the P1 actor holds (TrainingRemoteActor.input / TrainingBattleAIActor.input)
#define FGM_ACTOR2_SHIFT 26

#define FG_MAX_RECORDING_INPUT_FRAME 18000  // BattleCore.cs:67

struct FgVec4 { uint32_t x, y, z, w; };

FG_HD uint32_t fg_bits(uint32_t v, int shift, int n) { return (v >> shift) & ((1u << n) - 1u); }

FG_HD uint32_t fg_f2u(float f) { union { float f; uint32_t u; } c; c.f = f; return c.u; }
FG_HD float fg_u2f(uint32_t u) { union { float f; uint32_t u; } c; c.u = u; return c.f; }

// ---- host-side expansion (fg_get_state / fg_set_state) ----
static const int FG_ACTION_IDS[FT_NUM_ACTIONS] = FT_ACTION_IDS_INIT;
static const uint32_t FG_ACTION_INFO_H[FT_NUM_ACTIONS] = FT_ACTION_INFO_INIT;

static inline int fg_action_index(int id) {
    for (int i = 0; i < FT_NUM_ACTIONS; i++) if (FG_ACTION_IDS[i] == id) return i;
    return -1;
}

static const uint16_t FG_DASH_STATE_INFO[FT_NUM_DASH_STATES] = FT_DASH_STATE_INFO_INIT;   // since | runlen << 4 | lastdir << 8

// Left / Right bit history (bit i = held i frames ago) equivalent to a dash-automaton state
static inline void fg_dash_state_to_history(uint32_t id, uint32_t *left, uint32_t *right) {
    *left = 0; *right = 0;
    if (id == 0 || id >= FT_NUM_DASH_STATES) return;
    const uint32_t info = FG_DASH_STATE_INFO[id], since = info & 15u, runlen = (info >> 4) & 15u, lastdir = info >> 8;
    const uint32_t last = runlen >= 9u ? 16u : since + runlen;        // "9 or more": fill what a 16-frame history shows
    for (uint32_t k = since; k < last && k < 16u; k++) {
        if (lastdir & 1u) *left |= 1u << k;
        if (lastdir & 2u) *right |= 1u << k;
    }
}
// ... and back: any 16-frame bit history reduces to (since, runlen, lastdir)
static inline uint32_t fg_history_to_dash_state(uint32_t left, uint32_t right) {
    const uint32_t any = (left | right) & 0xffffu;
    uint32_t since = 0;
    while (since < 8u && !((any >> since) & 1u)) since++;
    if (since >= 8u) return 0u;
    uint32_t runlen = 0;
    while (since + runlen < 16u && ((any >> (since + runlen)) & 1u) && runlen < 9u) runlen++;
    const uint32_t lastdir = ((left >> since) & 1u) | ((right >> since) & 1u) << 1;
    const uint32_t want = since | runlen << 4 | lastdir << 8;
    for (uint32_t i = 1; i < FT_NUM_DASH_STATES; i++) if (FG_DASH_STATE_INFO[i] == want) return i;
    return 0u;
}

static inline void fg_decode_fighter(const FgVec4 &v, fg_fighter_state *o) {
    o->pos_x = fg_u2f(v.x);
    o->velocity_x = fg_u2f(v.y);
    uint32_t p = v.z;
    o->action_id = FG_ACTION_IDS[fg_bits(p, FGP_ACT_SHIFT, 5) % FT_NUM_ACTIONS];
    o->action_frame = (int)fg_bits(p, FGP_FRAME_SHIFT, 6);
    o->hitstun = (int)fg_bits(p, FGP_STUN_SHIFT, 5);
    o->guard = (int)fg_bits(p, FGP_GUARD_SHIFT, 2);
    o->vital = (int)fg_bits(p, FGP_VITAL_SHIFT, 1);
    o->hit_count = (int)fg_bits(p, FGP_HITCNT_SHIFT, 1);
    o->buffer_id = fg_bits(p, FGP_BUF_SHIFT, 1) ? 110 : -1;
    o->reserve_id = fg_bits(p, FGP_RSV_SHIFT, 1) ? 310 : -1;
    o->is_input_backward = (int)fg_bits(p, FGP_INBACK_SHIFT, 1);
    o->is_reserve_prox = (int)fg_bits(p, FGP_RPROX_SHIFT, 1);
    o->shake = (int)fg_bits(p, FGP_SHAKE_SHIFT, 3) * (fg_bits(p, FGP_SHAKE_SIGN_SHIFT, 1) ? -1 : 1);
    o->has_won = 0;
    fg_dash_state_to_history(fg_bits(v.w, FGH_DASH_SHIFT, 8), &o->hist_left, &o->hist_right);
    o->attack_run = (int)fg_bits(v.w, FGH_ARUN_SHIFT, 6);
    o->input0 = (int)((o->hist_left & 1u) | (o->hist_right & 1u) << 1 | (o->attack_run > 0 ? 4u : 0u));
}

// returns 0 on success, -1 if the state cannot be represented
static inline int fg_encode_fighter(const fg_fighter_state *s, FgVec4 *v) {
    int idx = fg_action_index(s->action_id);
    if (idx < 0 || idx == FT_IDX_WIN || s->has_won) return -1;
    int frame_count = (int)(FG_ACTION_INFO_H[idx] & 0x1ffu);
    int frame = s->action_frame;
    // at a frame boundary an action is never at its end (the frame that reaches frameCount switches action at once,
    // Fighter.cs:201-286); only a dead fighter's DEAD action keeps counting, and nothing reads it any more
    if (idx == FT_IDX_DEAD && frame > FGP_MAX_FRAME) frame = FGP_MAX_FRAME;
    if (frame < 0 || frame >= frame_count || frame > FGP_MAX_FRAME) return -1;
    if (s->hitstun < 0 || s->hitstun > 31 || s->guard < 0 || s->guard > 3 || s->vital < 0 || s->vital > 1) return -1;
    if (s->hit_count < 0 || s->hit_count > 1) return -1;
    if (!(s->buffer_id == -1 || s->buffer_id == 110) || !(s->reserve_id == -1 || s->reserve_id == 310)) return -1;
    if (s->shake < -6 || s->shake > 6 || s->attack_run < 0 || s->attack_run > 59) return -1;
    if ((s->hist_left | s->hist_right) >> 16) return -1;
    uint32_t carry = ((frame + 1 >= frame_count && s->hitstun == 0) ? FGP_CARRY_END : 0u) | ((FG_ACTION_INFO_H[idx] >> 9) & 1u ? FGP_CARRY_ALWAYS : 0u)
                   | ((idx == FT_IDX_N_ATTACK || idx == FT_IDX_B_ATTACK) ? FGP_CARRY_NORMAL : 0u);
    uint32_t mag = (uint32_t)(s->shake < 0 ? -s->shake : s->shake);
    uint32_t p = (uint32_t)idx << FGP_ACT_SHIFT | (uint32_t)frame << FGP_FRAME_SHIFT
               | (uint32_t)s->hitstun << FGP_STUN_SHIFT | (uint32_t)s->guard << FGP_GUARD_SHIFT
               | (uint32_t)s->vital << FGP_VITAL_SHIFT | (uint32_t)s->hit_count << FGP_HITCNT_SHIFT
               | (uint32_t)(s->buffer_id == 110) << FGP_BUF_SHIFT | (uint32_t)(s->reserve_id == 310) << FGP_RSV_SHIFT
               | (uint32_t)(s->is_input_backward != 0) << FGP_INBACK_SHIFT
               | (uint32_t)(s->is_reserve_prox != 0) << FGP_RPROX_SHIFT
               | mag << FGP_SHAKE_SHIFT | (uint32_t)(s->shake < 0) << FGP_SHAKE_SIGN_SHIFT | carry;
    v->x = fg_f2u(s->pos_x);
    v->y = fg_f2u(s->velocity_x);
    v->z = p;
    v->w = fg_history_to_dash_state(s->hist_left, s->hist_right) << FGH_DASH_SHIFT | (uint32_t)s->attack_run << FGH_ARUN_SHIFT;
    return 0;
}

static inline void fg_decode_env(const FgVec4 &f1, const FgVec4 &f2, const FgVec4 &e, const FgVec4 &r, fg_env_state *o) {
    uint32_t m = e.y;
    fg_decode_fighter(f1, &o->f[0]);
    fg_decode_fighter(f2, &o->f[1]);
    o->frame = (int32_t)e.x;
    o->recorded_input[0] = (int)fg_bits(m, FGM_REC1_SHIFT, 3);
    o->recorded_input[1] = (int)fg_bits(m, FGM_REC2_SHIFT, 3);
    o->done = (int)fg_bits(m, FGM_DONE_SHIFT, 1);
    o->cum_reward_index = (int)fg_bits(m, FGM_CUM_SHIFT, 4);
    o->actor_input[0] = (int)fg_bits(m, FGM_ACTOR1_SHIFT, 3);
    o->actor_input[1] = (int)fg_bits(m, FGM_ACTOR2_SHIFT, 3);
    o->rng_state[0] = r.x; o->rng_state[1] = r.y; o->rng_state[2] = r.z; o->rng_state[3] = r.w;
    o->bot_queue[0] = e.w; o->bot_queue[1] = e.z;
    o->p1_bot_memory = (int)fg_bits(m, FGM_P1MEM_SHIFT, 9);
}

static inline int fg_encode_env(const fg_env_state *s, FgVec4 *f1, FgVec4 *f2, FgVec4 *e, FgVec4 *r) {
    if (fg_encode_fighter(&s->f[0], f1) || fg_encode_fighter(&s->f[1], f2)) return -1;
    if (s->cum_reward_index < 0 || s->cum_reward_index >= FT_NUM_CUM) return -1;
    uint32_t m = ((uint32_t)s->recorded_input[0] & 7u) << FGM_REC1_SHIFT | ((uint32_t)s->recorded_input[1] & 7u) << FGM_REC2_SHIFT
               | (uint32_t)(s->done != 0) << FGM_DONE_SHIFT | (uint32_t)s->cum_reward_index << FGM_CUM_SHIFT
               | ((uint32_t)s->actor_input[0] & 7u) << FGM_ACTOR1_SHIFT | ((uint32_t)s->actor_input[1] & 7u) << FGM_ACTOR2_SHIFT
               | ((uint32_t)s->p1_bot_memory & 0x1ffu) << FGM_P1MEM_SHIFT;
    e->x = (uint32_t)s->frame; e->y = m; e->z = s->bot_queue[1]; e->w = s->bot_queue[0];
    r->x = s->rng_state[0]; r->y = s->rng_state[1]; r->z = s->rng_state[2]; r->w = s->rng_state[3];
    return 0;
}
#endif
