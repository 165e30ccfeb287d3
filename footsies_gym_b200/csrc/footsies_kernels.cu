// footsies_kernels.cu -- sm_100a kernels + C ABI of the batched FOOTSIES simulator.
//
// One CUDA thread owns one battle.  The whole reference frame update
//   BattleCore.FixedUpdate/UpdateFightState (BattleCore.cs:201-220, 347-364)
//   -> Fighter.UpdateInput / IncrementActionFrame / UpdateActionRequest / UpdateMovement / UpdateBoxes
//      (Fighter.cs:140-324, 472-510, 546-635, 671-719)
//   -> push / wall clamp / hitbox-hurtbox collision + damage (BattleCore.cs:483-591, Fighter.cs:352-454)
//   -> in-game bot (BattleAI.cs:41-403, queried as TrainingManager.cs:59-77 does)
//   -> observation, info, reward, termination (footsies.py:336-405, 518-570)
// runs in registers between one 64-byte state load and one 64-byte state store per env (four 16-byte
// SoA planes, fully coalesced 128-bit accesses).  Frame data is pre-expanded per (action, frame)
// (frame_tables.h) and staged once per CTA into shared memory; CTAs are persistent (grid-stride).
// No tensor cores: nothing here is a contraction.  The bound is HBM bandwidth (K = 1) or issue slots (K > 1).
//
// fp32 discipline: compiled with -fmad=false; every add/mul below rounds on its own exactly like the
// scalar C# expression it restates.  Multiplications by the facing sign (+-1) and by 0.5 are exact.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <new>
#include <vector>

#include "state_codec.h"

namespace {

#ifndef FG_THREADS
#define FG_THREADS 256
#endif
#ifndef FG_STAGES
#define FG_STAGES 2
#endif
#ifndef FG_BLOCKS_PER_SM
#define FG_BLOCKS_PER_SM 4
#endif
constexpr int kThreads = FG_THREADS;
constexpr uint32_t kFull = 0xffffffffu;

// action indices (moves.py order)
enum : uint32_t { STAND = FT_IDX_STAND, FORWARD = FT_IDX_FORWARD, BACKWARD = FT_IDX_BACKWARD,
                  DASH_FORWARD = FT_IDX_DASH_FORWARD, DASH_BACKWARD = FT_IDX_DASH_BACKWARD,
                  N_ATTACK = FT_IDX_N_ATTACK, B_ATTACK = FT_IDX_B_ATTACK, N_SPECIAL = FT_IDX_N_SPECIAL,
                  B_SPECIAL = FT_IDX_B_SPECIAL, DAMAGE = FT_IDX_DAMAGE, GUARD_BREAK = FT_IDX_GUARD_BREAK,
                  GUARD_PROXIMITY = FT_IDX_GUARD_PROXIMITY, DEAD = FT_IDX_DEAD, WIN = FT_IDX_WIN };

// ---- bot input patterns (BattleAI.cs:192-342); move values: 0 none, 1 forward, 2 backward ----
// move pattern ids: 1 Neutral, 2 FarApproach1, 3 FarApproach2, 4 MidApproach1, 5 MidApproach2, 6 FallBack1, 7 FallBack2
// attack pattern ids: 1 NoAttack, 2 OneHitImmediate, 3 TwoHitImmediate, 4 ImmediateSpecial, 5 DelaySpecial
constexpr int kMovePatBytes = 408;
constexpr int kAttPatBytes = 256;

struct __align__(16) Tables {
    uint4 rows[FT_NUM_ROWS];
    uint4 hit[8];
    uint2 hurt[16];
    uint2 push[8];
    uint32_t action_info[32];
    uint32_t attack[8];
    double term_reward[FT_NUM_CUM][4][2];
    double step_reward[4];
    uint8_t cum_next[16][4];
    uint32_t move_meta[8], att_meta[8];   // pattern offset | length << 16
    unsigned long long mod_magic[8];      // floor(2^64 / n) + 1
    uint8_t move_pat[kMovePatBytes];
    uint8_t att_pat[kAttPatBytes];
};
static_assert(sizeof(Tables) % 16 == 0, "Tables is copied as uint4");

struct Params {
    uint4 *pl_f1, *pl_f2, *pl_env, *pl_rng;
    unsigned long long *stats;
    const uint8_t *act1, *act2;
    float4 *obs;
    float *reward;
    uint8_t *terminated;
    int32_t *info_frame;
    uchar4 *info_misc;
    const Tables *tables;
    const uint8_t *mask;   // reset / seed kernels
    const uint8_t *step_mask;
    long long seed_base, first_env_index;
    int n, frame_skip, autoreset, stale_intro;
};

// Per-thread packed statistics: three 32-bit words of four 8-bit lanes each, bumped once per frame and folded
// by a warp reduction when a thread has gone kStatFlushFrames frames without a flush (and at kernel end).
//   A: episodes | P1 wins | P2 wins | double KOs      (1 << 8*w trick: w = winner code)
//   R: (none)   | hits    | blocks  | guard breaks    (1 << 8*DamageResult for each of the two attack passes)
//   S: specials | specials-from-neutral | resets | (unused)
struct StatAcc { uint32_t a, r, s; };
constexpr uint32_t kStatFlushFrames = 120u;   // <= 2 events per lane per frame -> a byte lane cannot overflow

__device__ __forceinline__ float u2f(uint32_t u) { return __uint_as_float(u); }
__device__ __forceinline__ uint32_t f2u(float f) { return __float_as_uint(f); }

struct Env {            // one battle, in registers
    float pos1, vel1, pos2, vel2;
    uint32_t pk1, hist1, pk2, hist2;
    int32_t frame;
    uint32_t misc, bq2, bq1;
    uint32_t r0, r1, r2, r3;
};

// UnityEngine.Random restated as xorshift128 (closed source; see DESIGN.md "parity unpinned")
__device__ __forceinline__ uint32_t rng_next(Env &e) {
    uint32_t t = e.r0 ^ (e.r0 << 11);
    e.r0 = e.r1; e.r1 = e.r2; e.r2 = e.r3;
    e.r3 = e.r3 ^ (e.r3 >> 19) ^ t ^ (t >> 8);
    return e.r3;
}

struct FrameOut {       // per-fighter products of the pre-collision phases
    uint32_t flags;     // expanded row flags (boxes present, hurt ids, push id)
    uint32_t kind;      // attack kind of the action the boxes were built from
    float pos_b;        // position when the boxes were built (after movement, before push)
};

// Fighter.UpdateInput + IncrementActionFrame + UpdateActionRequest + UpdateMovement for one fighter.
// SIDE 0 = P1 (faces right: forward = Right), 1 = P2 (faces left: forward = Left).
template <int SIDE>
__device__ __forceinline__ void update_fighter(const Tables &T, uint32_t in, float &pos, float &vel, uint32_t &pk,
                                               uint32_t &hist, uint32_t &arun, FrameOut &fo) {
    // ---- UpdateInput (Fighter.cs:172-188) on the compact history ----
    const uint32_t inA = (in >> 2) & 1u;
    const uint32_t hl = ((hist & 0xffffu) << 1) | (in & 1u);          // bit i = Left held i frames ago, i = 0..16
    const uint32_t hr = ((hist >> 16) << 1) | ((in >> 1) & 1u);
    const bool prev_a = arun != 0u;
    const bool atk_down = inA && !prev_a;                              // IsAttackInput(inputDown[0])
    const bool special = !inA && arun >= 59u;                          // CheckSpecialAttackInput (Fighter.cs:569-583)
    arun = inA ? min(arun + 1u, 59u) : 0u;
    hist = (hl & 0xffffu) | (hr << 16);
    const uint32_t fm = SIDE == 0 ? hr : hl;
    const uint32_t bm = SIDE == 0 ? hl : hr;
    const bool back = bm & 1u;
    // CheckForwardDashInput / CheckBackwardDashInput (Fighter.cs:585-635): first older frame (1..8) with a
    // direction held decides; then any neutral frame among the 8 frames before it.
    const uint32_t either = fm | bm;
    const uint32_t scan = either & 0x1feu;
    const int i = __ffs(scan | 0x200u) - 1;                            // 1..9 (9 = none)
    const bool gap = ((~either >> (i + 1)) & 0xffu) != 0u;
    const bool dash_f = (fm & 3u) == 1u && scan != 0u && !((bm >> i) & 1u) && gap;
    const bool dash_b = (bm & 3u) == 1u && scan != 0u && !((fm >> i) & 1u) && gap;

    // All fields are updated in place in the packed word (no unpack / repack).
    constexpr uint32_t M_STUN = 31u << FGP_STUN_SHIFT, M_GV = 7u << FGP_GUARD_SHIFT, M_HIT = 1u << FGP_HITCNT_SHIFT,
                       M_BUF = 1u << FGP_BUF_SHIFT, M_RSV = 1u << FGP_RSV_SHIFT, M_INBACK = 1u << FGP_INBACK_SHIFT,
                       M_RPROX = 1u << FGP_RPROX_SHIFT, M_SHAKE = 15u << FGP_SHAKE_SHIFT;
    // ---- IncrementActionFrame (Fighter.cs:140-166) ----
    if (pk & M_SHAKE) {                                                 // sprite shake decay (only after a hit)
        int shake = ((int)(pk << 1)) >> 28;
        shake = -shake; shake += shake > 0 ? -1 : 1;
        pk = (pk & ~M_SHAKE) | ((uint32_t)shake & 15u) << FGP_SHAKE_SHIFT;
    }
    pk = (pk & M_STUN) ? pk - (1u << FGP_STUN_SHIFT) : pk + (1u << FGP_FRAME_SHIFT);   // stun-- else frame++
    const bool stun0 = !(pk & M_STUN);
    const uint32_t act = pk & 31u;
    const uint32_t frame = (pk >> FGP_FRAME_SHIFT) & 511u;

    // ---- UpdateActionRequest (Fighter.cs:201-286) with the RequestAction chain (Fighter.cs:472-510) collapsed:
    //      when the action ended or is alwaysCancelable the FIRST request of the chain wins, otherwise the only
    //      effect a request can have is buffering N_SPECIAL inside a cancel window. ----
    const uint32_t info0 = T.action_info[act];
    const bool ended = frame >= (info0 & 511u);
    bool want_buffer = false;
    bool set;
    uint32_t req;
    if ((pk & (M_RSV | M_STUN)) == M_RSV) {                             // reserved GUARD_BREAK (Fighter.cs:212-218)
        req = GUARD_BREAK; set = true;
    } else if ((pk & (M_BUF | M_HIT | M_STUN)) == (M_BUF | M_HIT)) {    // buffered cancel (Fighter.cs:222-229)
        req = N_SPECIAL; set = true;
    } else {
        const uint32_t dir = (fm | bm) & 1u;
        const bool in_normal = (act == N_ATTACK || act == B_ATTACK) && !ended;
        // attack request: N_ATTACK 5 / B_ATTACK 6 / N_SPECIAL 7 / B_SPECIAL 8 (the B_ variant when a direction is held)
        const uint32_t areq = special ? N_SPECIAL + dir : in_normal ? (uint32_t)N_SPECIAL : N_ATTACK + dir;
        // movement request by (fwd, back, isReserveProximityGuard): STAND / FORWARD / BACKWARD / GUARD_PROXIMITY
        // nibble LUT indexed by the raw Left/Right bits (+4 when the proximity-guard flag is set)
        const uint32_t mv = ((SIDE == 0 ? 0x01E00120u : 0x0E100210u) >> (4u * ((in & 3u) | ((pk >> (FGP_RPROX_SHIFT - 2)) & 4u)))) & 15u;
        req = (special || atk_down) ? areq : dash_f ? (uint32_t)DASH_FORWARD : dash_b ? (uint32_t)DASH_BACKWARD : mv;
        const bool free_to_switch = ended || (info0 & 512u);
        set = free_to_switch && (ended || req != act);
        want_buffer = !free_to_switch && req == N_SPECIAL;
        pk = (pk & ~(M_INBACK | M_RPROX)) | (back ? M_INBACK : 0u);     // isInputBackward = back; reserve flag consumed
    }
    if (set) pk = (pk & (M_STUN | M_GV | M_INBACK | M_RPROX)) | req;    // SetCurrentAction (Fighter.cs:546-563)

    // ---- frame data of the (action, frame) the fighter ends up in ----
    const uint32_t info = T.action_info[pk & 31u];
    const uint32_t row_idx = ((info >> 14) & 1023u) + min((pk >> FGP_FRAME_SHIFT) & 511u, (info >> 24) & 63u);
    const uint4 row = T.rows[row_idx];
    if (want_buffer && (row.z & 8u)) pk |= M_BUF;                       // cancel window (Fighter.cs:492-505)

    // ---- UpdateMovement (Fighter.cs:291-319) ----
    if (stun0) {
        const float dx = u2f(row.x);
        pos = pos + (SIDE == 0 ? dx : -dx);
        if (row.z & 1u) vel = u2f(row.y);
    }
    fo.flags = row.z;
    fo.kind = (info >> 11) & 7u;
    fo.pos_b = pos;
}

// World x-extent of a box built at position pos_b (Fighter.cs:706-719: x = pos + data.x * sign; BoxBase xMin/xMax,
// Fighter.cs:12-13) and then displaced by the push (s) and the wall clamp (t) like ApplyPositionChange does to
// already-built boxes (Fighter.cs:331-350).  Every operation rounds separately.
template <int SIDE>
__device__ __forceinline__ void box_extent(float pos_b, uint32_t cx_bits, uint32_t hw_bits, float s, float t,
                                           float &lo, float &hi) {
    const float cx = u2f(cx_bits), hw = u2f(hw_bits);
    const float x = ((pos_b + (SIDE == 0 ? cx : -cx)) + s) + t;
    lo = x - hw;
    hi = x + hw;
}

// Geometry half of BattleCore.UpdateHitboxHurtboxCollision (BattleCore.cs:535-565) for one attacker: does its real /
// proximity hitbox overlap any of the victim's (<= 2) hurtboxes?  BoxBase.Overlaps (Fighter.cs:17-25, inclusive); the
// y half of each test is pre-resolved into the hitbox's mask over hurtbox ids.  Straight-line code: the boxes are a
// snapshot, so both attackers' tests can be evaluated before either attack is applied.
template <int ASIDE>
__device__ __forceinline__ void attack_overlaps(const Tables &T, const FrameOut &af, const FrameOut &vf, float a_s, float a_t,
                                                float v_s, float v_t, bool &real_hit, bool &prox_hit) {
    const uint32_t kidx = (max(af.kind, 1u) - 1u) * 2u;
    const uint4 hp = T.hit[kidx], hr = T.hit[kidx + 1u];               // proximity box, real box of this attack
    const uint32_t id0 = (vf.flags >> 4) & 15u, id1 = (vf.flags >> 8) & 15u;
    const uint2 v0 = T.hurt[id0], v1 = T.hurt[id1];
    float plo, phi, rlo, rhi, v0lo, v0hi, v1lo, v1hi;
    box_extent<ASIDE>(af.pos_b, hp.x, hp.y, a_s, a_t, plo, phi);
    box_extent<ASIDE>(af.pos_b, hr.x, hr.y, a_s, a_t, rlo, rhi);
    box_extent<1 - ASIDE>(vf.pos_b, v0.x, v0.y, v_s, v_t, v0lo, v0hi);
    box_extent<1 - ASIDE>(vf.pos_b, v1.x, v1.y, v_s, v_t, v1lo, v1hi);
    // otherBox.xMax >= xMin && otherBox.xMin <= xMax with self = hitbox, other = hurtbox; id 0 (no box) never has its bit set
    const bool r0 = ((hr.z >> id0) & 1u) && v0hi >= rlo && v0lo <= rhi;
    const bool r1 = ((hr.z >> id1) & 1u) && v1hi >= rlo && v1lo <= rhi;
    const bool p0 = ((hp.z >> id0) & 1u) && v0hi >= plo && v0lo <= phi;
    const bool p1 = ((hp.z >> id1) & 1u) && v1hi >= plo && v1lo <= phi;
    real_hit = (af.flags & 4u) && (r0 || r1);
    prox_hit = (af.flags & 2u) && (p0 || p1);
}

// Effect half of one attacker -> victim pass (BattleCore.cs:567-586): NotifyAttackHit / NotifyDamaged /
// GetHitStunFrame / SetHitStun / SetSpriteShakeFrame / NotifyInProximityGuardRange (Fighter.cs:352-454).
// Hit counts and the victim's action are the CURRENT ones (P1's hit may just have changed P2), boxes are the snapshot.
template <int ASIDE>
__device__ __forceinline__ uint32_t attack_apply(const Tables &T, uint32_t &apk, uint32_t &vpk, const FrameOut &af,
                                                 bool real_hit, bool prox_hit) {
    const bool can = (af.flags & 6u) && !((apk >> FGP_HITCNT_SHIFT) & 1u);   // a hitbox is out and CanAttackHit
    if (can && real_hit) {
        const uint32_t atk = T.attack[af.kind];
        uint32_t guard = (vpk >> FGP_GUARD_SHIFT) & 3u;
        const bool brk = guard == 0u;                                   // guardHealth < 0 after the decrement
        guard = brk ? 0u : guard - 1u;
        uint32_t vital = (vpk >> FGP_VITAL_SHIFT) & 1u;
        const uint32_t vact = vpk & 31u;
        const bool guarding = vact == BACKWARD || ((T.action_info[vact] >> 10) & 1u);
        uint32_t nact, stun, rsv = 0u, res;
        if (guarding) {
            nact = (atk >> 5) & 31u;
            rsv = brk ? 1u : 0u;
            stun = brk ? (atk >> 21) & 31u : (atk >> 16) & 31u;
            res = brk ? 3u : 2u;
        } else {
            if ((atk >> 10) & 1u) vital = 0u;
            nact = atk & 31u;
            stun = (atk >> 11) & 31u;
            res = 1u;
        }
        const int sh = min((int)stun / 3, 6) * (ASIDE == 0 ? 1 : -1);   // victim of P1 faces left -> +
        vpk = nact | stun << FGP_STUN_SHIFT | guard << FGP_GUARD_SHIFT | vital << FGP_VITAL_SHIFT
            | rsv << FGP_RSV_SHIFT | (vpk & (3u << FGP_INBACK_SHIFT)) | ((uint32_t)sh & 15u) << FGP_SHAKE_SHIFT;
        apk = (apk & ~(31u << FGP_STUN_SHIFT)) | stun << FGP_STUN_SHIFT | 1u << FGP_HITCNT_SHIFT;
        return res;
    }
    // NotifyInProximityGuardRange: latch only while the victim holds back (Fighter.cs:400-406)
    if (can && prox_hit) vpk |= ((vpk >> FGP_INBACK_SHIFT) & 1u) << FGP_RPROX_SHIFT;
    return 0u;
}

// BattleAI.getNextAIInput (BattleAI.cs:41-66) on pattern-id + cursor queues.  `dist` and `opp_act` are the state
// captured by the PREVIOUS call (the ascending shift loop at BattleAI.cs:358-361 makes fightStates[5] exactly that).
// r % n for 2 <= n <= 7 without a division: floor(r / n) == umul64hi(r, floor(2^64 / n) + 1) for every 32-bit r.
__device__ __forceinline__ uint32_t mod_small(const Tables &T, uint32_t r, uint32_t n) {
    const uint32_t q = (uint32_t)__umul64hi((unsigned long long)r, T.mod_magic[n]);
    return r - q * n;
}

template <int SIDE>
__device__ __forceinline__ uint32_t bot_next(const Tables &T, Env &e, uint32_t &q, float dist, uint32_t opp_act) {
    // queue word: move position in move_pat [0:9) | moves remaining [9:16) | attack position in att_pat [16:24) |
    // attacks remaining [24:31): dequeuing is one add on the packed word
    uint32_t input = 0u;
    const bool have_m = (q & (127u << 9)) != 0u, have_a = (q & (127u << 24)) != 0u;
    if (have_m) {
        const uint32_t v = T.move_pat[q & 511u];                        // 0 none, 1 forward, 2 backward
        q += 1u - (1u << 9);
        input = SIDE == 1 ? v : ((v >> 1) | ((v & 1u) << 1));           // P2: forward = Left(1); P1: forward = Right(2)
    }
    if (have_a) {
        input |= T.att_pat[(q >> 16) & 255u];
        q += (1u << 16) - (1u << 24);
    }
    if (!(have_m && have_a)) {                                          // an empty queue is refilled and contributes 0 (BattleAI.cs:50-62)
        const int bucket = dist > 4.0f ? 0 : dist > 3.0f ? 1 : dist > 2.5f ? 2 : dist > 2.0f ? 3 : 4;
        if (!have_m) {                                                  // SelectMovement (BattleAI.cs:68-126)
            const uint32_t n = (0x34572u >> (4 * bucket)) & 15u;        // Random.Range(0, n): n = 2,7,5,4,3
            const uint32_t r = mod_small(T, rng_next(e), n);
            // nibble r of the bucket's word = move pattern id
            const uint32_t sel = bucket == 0 ? 0x32u : bucket == 1 ? 0x1325544u : bucket == 2 ? 0x17654u
                               : bucket == 3 ? 0x1176u : 0x176u;
            const uint32_t meta = T.move_meta[(sel >> (4 * r)) & 15u];  // offset | length << 16
            q = (q & 0xffff0000u) | (meta & 0xffffu) | (meta >> 16) << 9;
        }
        if (!have_a) {                                                  // SelectAttack (BattleAI.cs:128-190)
            const bool opp_hurt = opp_act == DAMAGE || opp_act == GUARD_BREAK || opp_act == N_SPECIAL || opp_act == B_SPECIAL;
            const bool opp_normal = opp_act == N_ATTACK || opp_act == B_ATTACK;
            uint32_t ap;
            if (opp_hurt || (bucket == 1 && opp_normal)) {
                ap = 3u;                                                // AddTwoHitImmediateAttack, no draw
            } else {
                const uint32_t n = (0x36354u >> (4 * bucket)) & 15u;    // n = 4,5,3,6,3
                const uint32_t r = mod_small(T, rng_next(e), n);
                const uint32_t sel = bucket == 0 ? 0x1111u : bucket == 1 ? 0x52211u : bucket == 2 ? 0x321u
                                   : bucket == 3 ? 0x543322u : 0x332u;
                ap = (sel >> (4 * r)) & 15u;
            }
            const uint32_t meta = T.att_meta[ap];
            q = (q & 0x0000ffffu) | (meta & 0xffffu) << 16 | (meta >> 16) << 24;
        }
    }
    return input;
}

// Stop -> Intro -> one Intro frame -> Fight (BattleCore.cs:176-200, 262-291, 329-345) for one env.
// What survives from the previous round (SetupBattleStart, Fighter.cs:120-135, does not touch them): the actors'
// held inputs (replayed by the Intro frame), hit stun, isInputBackward / isReserveProximityGuard.
template <bool P1BOT, bool P2BOT>
__device__ __forceinline__ void reset_env(const Tables &T, Env &e, bool stale_intro) {
    const bool was_done = (e.misc >> FGM_DONE_SHIFT) & 1u;
    uint32_t a1 = (e.misc >> FGM_ACTOR1_SHIFT) & 7u, a2 = (e.misc >> FGM_ACTOR2_SHIFT) & 7u;
    if (!stale_intro) { a1 = 0u; a2 = 0u; }
    uint32_t pk[2] = { e.pk1, e.pk2 };
    uint32_t npk[2];
#pragma unroll
    for (int s = 0; s < 2; s++) {
        uint32_t stun = (pk[s] >> FGP_STUN_SHIFT) & 31u;
        uint32_t keep = pk[s] & (3u << FGP_INBACK_SHIFT);
        if (was_done) {
            // the End-state frame (BattleCore.cs:371-381) ran once: hit stun ticks; a dead fighter went through the
            // normal request path with cleared inputs (flags reset), a winner returned early (flags kept)
            if (stun > 0u) stun--;
            if (!((pk[s] >> FGP_VITAL_SHIFT) & 1u)) keep = 0u;
        }
        // Intro frame: IncrementActionFrame (frame 0 -> 1 unless in hit stun), RequestAction(STAND) is a no-op
        uint32_t frame = 1u;
        if (stun > 0u) { stun--; frame = 0u; }
        npk[s] = STAND | frame << FGP_FRAME_SHIFT | stun << FGP_STUN_SHIFT | 3u << FGP_GUARD_SHIFT
               | 1u << FGP_VITAL_SHIFT | keep;
    }
    e.pk1 = npk[0]; e.pk2 = npk[1];
    e.pos1 = -2.0f; e.pos2 = 2.0f; e.vel1 = 0.0f; e.vel2 = 0.0f;
    e.hist1 = (a1 & 1u) | ((a1 >> 1) & 1u) << 16;                       // UpdateInput(stale) after ClearInput
    e.hist2 = (a2 & 1u) | ((a2 >> 1) & 1u) << 16;
    e.frame = -1;
    e.bq1 = 0u; e.bq2 = 0u;                                            // BattleAI.Reset (BattleAI.cs:393-403)
    const uint32_t run1 = (a1 >> 2) & 1u, run2 = (a2 >> 2) & 1u;       // Attack run after the Intro frame's input
    // first bot query at the Fight transition (BattleCore.cs:289): decision input = round-start state
    if (P1BOT) a1 = bot_next<0>(T, e, e.bq1, 4.0f, STAND);
    if (P2BOT) a2 = bot_next<1>(T, e, e.bq2, 4.0f, STAND);
    e.misc = run1 << FGM_ARUN1_SHIFT | run2 << FGM_ARUN2_SHIFT
           | a1 << FGM_ACTOR1_SHIFT | a2 << FGM_ACTOR2_SHIFT;           // recorded inputs 0, done 0, cum 0
}

// FootsiesEnv._extract_obs / _extract_info (footsies.py:336-380) incl. the DEAD/WIN -> STAND remap of step()
// (footsies.py:538-549; a no-op on the reset observation, which is always STAND).
__device__ __forceinline__ void write_outputs(const Params &p, int i, const Env &e, float reward, bool terminated) {
    uint32_t m1 = e.pk1 & 31u, m2 = e.pk2 & 31u;
    if (m1 >= DEAD) m1 = STAND;
    if (m2 >= DEAD) m2 = STAND;
    const uint32_t f1 = m1 <= BACKWARD ? 0u : (e.pk1 >> FGP_FRAME_SHIFT) & 511u;
    const uint32_t f2 = m2 <= BACKWARD ? 0u : (e.pk2 >> FGP_FRAME_SHIFT) & 511u;
    float4 o0, o1;
    o0.x = (float)((e.pk1 >> FGP_GUARD_SHIFT) & 3u); o0.y = (float)((e.pk2 >> FGP_GUARD_SHIFT) & 3u);
    o0.z = (float)m1; o0.w = (float)m2;
    o1.x = (float)f1; o1.y = (float)f2; o1.z = e.pos1; o1.w = e.pos2;
    p.obs[2 * (size_t)i] = o0;
    p.obs[2 * (size_t)i + 1] = o1;
    p.reward[i] = reward;
    p.terminated[i] = terminated ? 1 : 0;
    p.info_frame[i] = e.frame;
    uchar4 im;
    im.x = (e.misc >> FGM_REC1_SHIFT) & 7u; im.y = (e.misc >> FGM_REC2_SHIFT) & 7u;
    im.z = (e.pk1 >> FGP_STUN_SHIFT) & 31u; im.w = (e.pk2 >> FGP_STUN_SHIFT) & 31u;
    p.info_misc[i] = im;
}

__device__ __forceinline__ void load_tables(Tables *dst, const Tables *src) {
    const uint4 *s = reinterpret_cast<const uint4 *>(src);
    uint4 *d = reinterpret_cast<uint4 *>(dst);
    for (int k = threadIdx.x; k < (int)(sizeof(Tables) / 16); k += blockDim.x) d[k] = s[k];
    __syncthreads();
}

template <bool WITH_RNG>
__device__ __forceinline__ void load_env(const Params &p, int i, Env &e) {
    const uint4 a = p.pl_f1[i], b = p.pl_f2[i], c = p.pl_env[i];
    e.pos1 = u2f(a.x); e.vel1 = u2f(a.y); e.pk1 = a.z; e.hist1 = a.w;
    e.pos2 = u2f(b.x); e.vel2 = u2f(b.y); e.pk2 = b.z; e.hist2 = b.w;
    e.frame = (int32_t)c.x; e.misc = c.y; e.bq2 = c.z; e.bq1 = c.w;
    if (WITH_RNG) { const uint4 r = p.pl_rng[i]; e.r0 = r.x; e.r1 = r.y; e.r2 = r.z; e.r3 = r.w; }
}
template <bool WITH_RNG>
__device__ __forceinline__ void store_env(const Params &p, int i, const Env &e) {
    p.pl_f1[i] = make_uint4(f2u(e.pos1), f2u(e.vel1), e.pk1, e.hist1);
    p.pl_f2[i] = make_uint4(f2u(e.pos2), f2u(e.vel2), e.pk2, e.hist2);
    p.pl_env[i] = make_uint4((uint32_t)e.frame, e.misc, e.bq2, e.bq1);
    if (WITH_RNG) p.pl_rng[i] = make_uint4(e.r0, e.r1, e.r2, e.r3);
}

// Warp-cooperative fold of the packed per-thread counters into the CTA's shared-memory vector.
__device__ __forceinline__ void flush_stats(StatAcc &acc, unsigned long long *s_stats, int lane) {
    const uint32_t w[3] = { acc.a, acc.r, acc.s };
    // (word, byte lane) -> statistic index; -1 = unused
    const int map[3][4] = { { FG_STAT_EPISODES, FG_STAT_P1_WINS, FG_STAT_P2_WINS, FG_STAT_DOUBLE_KO },
                            { -1, FG_STAT_HITS, FG_STAT_BLOCKS, FG_STAT_GUARD_BREAKS },
                            { FG_STAT_P1_SPECIALS, FG_STAT_P1_SPECIALS_NEUTRAL, FG_STAT_RESETS, -1 } };
#pragma unroll
    for (int k = 0; k < 3; k++)
#pragma unroll
        for (int b = 0; b < 4; b++) {
            if (map[k][b] < 0) continue;
            const uint32_t tot = __reduce_add_sync(kFull, (w[k] >> (8 * b)) & 255u);
            if (lane == 0 && tot) atomicAdd(&s_stats[map[k][b]], (unsigned long long)tot);
        }
    acc.a = 0u; acc.r = 0u; acc.s = 0u;
}

// One fight frame for one env (everything between "inputs known" and "state after the frame").
// Sets `terminal`, accumulates the Python float64 reward into `reward`, bumps the packed statistics.
template <bool P1BOT, bool P2BOT, bool DENSE>
__device__ __forceinline__ void simulate_frame(const Tables &T, Env &e, uint32_t in1, uint32_t in2, double &reward,
                                               bool &terminal, StatAcc &acc, unsigned long long *s_stats) {
    // state the bots will be shown after this frame (previous call's capture == state before this frame)
    const float pre_dist = fabsf(e.pos2 - e.pos1);
    const uint32_t pre_a1 = e.pk1 & 31u, pre_a2 = e.pk2 & 31u;
    const uint32_t g1_before = (e.pk1 >> FGP_GUARD_SHIFT) & 3u, g2_before = (e.pk2 >> FGP_GUARD_SHIFT) & 3u;

    e.frame++;
    // BattleCore.RecordInput (BattleCore.cs:593-607): recording stops after maxRecordingInputFrame frames
    if (e.frame < FG_MAX_RECORDING_INPUT_FRAME)
        e.misc = (e.misc & ~(63u << FGM_REC1_SHIFT)) | in1 << FGM_REC1_SHIFT | in2 << FGM_REC2_SHIFT;

    uint32_t arun1 = (e.misc >> FGM_ARUN1_SHIFT) & 63u, arun2 = (e.misc >> FGM_ARUN2_SHIFT) & 63u;
    FrameOut f1, f2;
    update_fighter<0>(T, in1, e.pos1, e.vel1, e.pk1, e.hist1, arun1, f1);
    update_fighter<1>(T, in2, e.pos2, e.vel2, e.pk2, e.hist2, arun2, f2);

    // ---- UpdatePushCharacterVsCharacter (BattleCore.cs:483-501), UnityEngine.Rect semantics: x = left edge, strict ----
    const uint2 pb1 = T.push[(f1.flags >> 12) & 7u], pb2 = T.push[(f2.flags >> 12) & 7u];
    const float px1 = e.pos1 + u2f(pb1.x), w1 = u2f(pb1.y);
    const float px2 = e.pos2 - u2f(pb2.x), w2 = u2f(pb2.y);
    const float xmax1 = w1 + px1, xmax2 = w2 + px2;
    float s1 = 0.0f, s2 = 0.0f;
    if (xmax2 > px1 && px2 < xmax1) {
        if (e.pos1 < e.pos2) { const float d = xmax1 - px2; s1 = -0.5f * d; s2 = 0.5f * d; }
        else if (e.pos1 > e.pos2) { const float d = xmax2 - px1; s1 = 0.5f * d; s2 = -0.5f * d; }
    }
    // ---- UpdatePushCharacterVsBackground (BattleCore.cs:503-519), BoxBase semantics: x = centre ----
    float t1 = 0.0f, t2 = 0.0f;
    {
        const float c = px1 + s1, hw = 0.5f * w1, mn = c - hw, mx = c + hw;
        if (mn < -5.0f) t1 = -5.0f - mn; else if (mx > 5.0f) t1 = 5.0f - mx;
    }
    {
        const float c = px2 + s2, hw = 0.5f * w2, mn = c - hw, mx = c + hw;
        if (mn < -5.0f) t2 = -5.0f - mn; else if (mx > 5.0f) t2 = 5.0f - mx;
    }
    e.pos1 = (e.pos1 + s1) + t1;
    e.pos2 = (e.pos2 + s2) + t2;

    // ---- UpdateHitboxHurtboxCollision (BattleCore.cs:521-591): P1 attacks first, then P2 with snapshot boxes ----
    uint32_t res_a = 0u, res_b = 0u;
    if ((f1.flags | f2.flags) & 6u) {                                   // somebody has a hitbox out
        bool real_a, prox_a, real_b, prox_b;
        attack_overlaps<0>(T, f1, f2, s1, t1, s2, t2, real_a, prox_a);
        attack_overlaps<1>(T, f2, f1, s2, t2, s1, t1, real_b, prox_b);
        res_a = attack_apply<0>(T, e.pk1, e.pk2, f1, real_a, prox_a);  // result on P2
        res_b = attack_apply<1>(T, e.pk2, e.pk1, f2, real_b, prox_b);  // result on P1
    }

    acc.r += (1u << (8u * res_a)) + (1u << (8u * res_b));                // byte lane = DamageResult of each pass
    const uint32_t a1 = e.pk1 & 31u;
    if (a1 != pre_a1 && (a1 - N_SPECIAL) < 2u)                          // became N_SPECIAL / B_SPECIAL (wrappers/statistics.py:36-46)
        acc.s += (pre_a1 != N_ATTACK && pre_a1 != B_ATTACK) ? 0x101u : 1u;

    // ---- KO (BattleCore.cs:212-217), termination (footsies.py:555) ----
    const bool dead1 = !((e.pk1 >> FGP_VITAL_SHIFT) & 1u), dead2 = !((e.pk2 >> FGP_VITAL_SHIFT) & 1u);
    terminal = dead1 || dead2;

    // ---- reward (footsies.py:382-405); Python floats are doubles ----
    if (DENSE) {
        const uint32_t code = (((e.pk1 >> FGP_GUARD_SHIFT) & 3u) < g1_before ? 1u : 0u)
                            | (((e.pk2 >> FGP_GUARD_SHIFT) & 3u) < g2_before ? 2u : 0u);
        uint32_t cum = (e.misc >> FGM_CUM_SHIFT) & 15u;
        if (code | (terminal ? 1u : 0u)) {
            cum = T.cum_next[cum][code];
            e.misc = (e.misc & ~(15u << FGM_CUM_SHIFT)) | cum << FGM_CUM_SHIFT;
            reward += terminal ? T.term_reward[cum][code][dead2 ? 1 : 0] : T.step_reward[code];
        }
    } else if (terminal) {
        reward += dead2 ? 1.0 : -1.0;
    }

    if (terminal) {
        // ChangeRoundState(KO): ClearInput on both fighters (BattleCore.cs:292-299); actors keep their inputs
        e.hist1 = 0u; e.hist2 = 0u; arun1 = 0u; arun2 = 0u;
        acc.a += 1u + (1u << (8u * (dead1 && dead2 ? 3u : dead2 ? 1u : 2u)));
        atomicAdd(&s_stats[FG_STAT_EPISODE_FRAMES], (unsigned long long)(e.frame + 1));   // rare: once per episode
        e.misc |= 1u << FGM_DONE_SHIFT;
    }
    // ---- TrainingManager.Step (TrainingManager.cs:59-77): actors' inputs for the next frame; bots are asked
    //      after the frame, P1 first, and not on the terminal frame ----
    uint32_t n1 = in1, n2 = in2;
    if (!terminal) {
        if (P1BOT) n1 = bot_next<0>(T, e, e.bq1, pre_dist, pre_a2);
        if (P2BOT) n2 = bot_next<1>(T, e, e.bq2, pre_dist, pre_a1);
    }
    e.misc = (e.misc & ~((63u << FGM_ARUN1_SHIFT) | (63u << FGM_ARUN2_SHIFT) | (63u << FGM_ACTOR1_SHIFT)))
           | arun1 << FGM_ARUN1_SHIFT | arun2 << FGM_ARUN2_SHIFT | n1 << FGM_ACTOR1_SHIFT | n2 << FGM_ACTOR2_SHIFT;
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
#ifndef FG_WAIT_MODE
#define FG_WAIT_MODE 0
#endif
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
#if FG_WAIT_MODE == 1
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)), "r"(parity), "r"(0x989680u) : "memory");
#elif FG_WAIT_MODE == 3
    uint32_t ok = 0u;
    while (true) {
        asm volatile(
            "{\n"
            ".reg .pred P1;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
            "selp.u32 %0, 1, 0, P1;\n"
            "}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (ok) break;
        __nanosleep(64);
    }
#else
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
#endif
}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

constexpr int kStages = FG_STAGES;
template <int PLANES> struct __align__(128) StageBuf { uint4 pl[PLANES][kThreads]; };

// FootsiesEnv.step for every env: up to K fused fight frames, or the reset of a finished env (autoreset).
// Persistent CTAs walk chunks of 256 consecutive envs; chunk c+grid is prefetched by one elected thread with
// 3-4 bulk copies of 4 KB (one per state plane) while chunk c is simulated out of registers.
template <bool KFUSED, bool P1BOT, bool P2BOT, bool DENSE>
__global__ void __launch_bounds__(kThreads) step_kernel(const Params p) {
    constexpr bool kRng = P1BOT || P2BOT;
    constexpr int kPlanes = kRng ? 4 : 3;
    __shared__ Tables T;
    __shared__ StageBuf<kPlanes> stage[kStages];
    __shared__ uint64_t full_bar[kStages], empty_bar[kStages];
    __shared__ unsigned long long s_stats[FG_STAT_COUNT];
    load_tables(&T, p.tables);
    if (threadIdx.x < FG_STAT_COUNT) s_stats[threadIdx.x] = 0ull;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; s++) { mbar_init(&full_bar[s], 1u); mbar_init(&empty_bar[s], kThreads / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int num_chunks = (p.n + kThreads - 1) / kThreads;
    const int full_chunks = p.n / kThreads;                             // chunks that can be bulk-copied whole
    const uint4 *const planes[4] = { p.pl_f1, p.pl_f2, p.pl_env, p.pl_rng };
    auto issue = [&](int chunk, int s) {                                // elected thread only
        mbar_expect_tx(&full_bar[s], (uint32_t)(kPlanes * kThreads * sizeof(uint4)));
#pragma unroll
        for (int k = 0; k < kPlanes; k++)
            tma_load_1d(stage[s].pl[k], planes[k] + (size_t)chunk * kThreads, (uint32_t)(kThreads * sizeof(uint4)), &full_bar[s]);
    };
    // prologue: chunks 0 .. kStages-2 of this CTA are in flight before the loop starts
    if (threadIdx.x == 0)
        for (int j = 0; j < kStages - 1; j++) {
            const int cj = blockIdx.x + j * gridDim.x;
            if (cj < full_chunks) issue(cj, j);
        }
    StatAcc acc = { 0u, 0u, 0u };
    uint32_t frames_done = 0u, frames_since_flush = 0u;
    // actions are prefetched one chunk ahead into registers
    uint32_t nin1 = 0u, nin2 = 0u;
    {
        const int i0 = blockIdx.x * kThreads + threadIdx.x;
        if (i0 < p.n) { if (!P1BOT) nin1 = p.act1[i0]; if (!P2BOT) nin2 = p.act2[i0]; }
    }
    int k = 0;
    for (int c = blockIdx.x; c < num_chunks; c += gridDim.x, k++) {
        const int s = k % kStages;
        const int i = c * kThreads + threadIdx.x;
        const bool valid = i < p.n && (p.step_mask == nullptr || p.step_mask[i < p.n ? i : 0] != 0);
        const bool staged = c < full_chunks;
#if FG_WAIT_MODE == 2
        // one warp polls the transaction barrier, the others block in bar.sync (no issue slots burnt); passing this
        // barrier also proves every warp has finished reading the stage of chunk k-1, so it can be refilled
        if (staged) {
            if (threadIdx.x < 32) mbar_wait(&full_bar[s], (k / kStages) & 1);
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            const int cn = c + (kStages - 1) * gridDim.x;
            if (cn < full_chunks) issue(cn, (k + kStages - 1) % kStages);
        }
#else
        if (threadIdx.x == 0) {                                         // producer: chunk k + kStages - 1 -> the stage read at k - 1
            const int cn = c + (kStages - 1) * gridDim.x;
            if (cn < full_chunks) {
                const int sn = (k + kStages - 1) % kStages;
                if (k >= 1) mbar_wait(&empty_bar[sn], ((k - 1) / kStages) & 1);   // every warp has read chunk k-1
                issue(cn, sn);
            }
        }
#endif
        const uint32_t act1 = nin1, act2 = nin2;
        {
            const int in = i + gridDim.x * kThreads;
            if (in < p.n) { if (!P1BOT) nin1 = p.act1[in]; if (!P2BOT) nin2 = p.act2[in]; }
        }
        Env e;
        bool run = false;
        uint32_t in1 = 0u, in2 = 0u;
        if (staged) {
#if FG_WAIT_MODE != 2
            mbar_wait(&full_bar[s], (k / kStages) & 1);
#endif
            const uint4 a = stage[s].pl[0][threadIdx.x], b = stage[s].pl[1][threadIdx.x], cc = stage[s].pl[2][threadIdx.x];
            e.pos1 = u2f(a.x); e.vel1 = u2f(a.y); e.pk1 = a.z; e.hist1 = a.w;
            e.pos2 = u2f(b.x); e.vel2 = u2f(b.y); e.pk2 = b.z; e.hist2 = b.w;
            e.frame = (int32_t)cc.x; e.misc = cc.y; e.bq2 = cc.z; e.bq1 = cc.w;
            if (kRng) { const uint4 r = stage[s].pl[kPlanes - 1][threadIdx.x]; e.r0 = r.x; e.r1 = r.y; e.r2 = r.z; e.r3 = r.w; }
#if FG_WAIT_MODE != 2
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[s]);
#endif
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
                simulate_frame<P1BOT, P2BOT, DENSE>(T, e, in1, in2, reward, terminal, acc, s_stats);
                frames_done++;
                if (KFUSED) {
                    if (P1BOT) in1 = (e.misc >> FGM_ACTOR1_SHIFT) & 7u;
                    if (P2BOT) in2 = (e.misc >> FGM_ACTOR2_SHIFT) & 7u;
                }
            }
        }
        if (run) {
            store_env<kRng>(p, i, e);
            write_outputs(p, i, e, (float)reward, terminal);
        }
        frames_since_flush += (uint32_t)K;                              // uniform across the CTA
        if (frames_since_flush >= kStatFlushFrames) { flush_stats(acc, s_stats, lane); frames_since_flush = 0u; }
    }
    flush_stats(acc, s_stats, lane);
    const uint32_t fsum = __reduce_add_sync(kFull, frames_done);
    if (lane == 0 && fsum) atomicAdd(&s_stats[FG_STAT_ENV_FRAMES], (unsigned long long)fsum);
    __syncthreads();
    if (threadIdx.x < FG_STAT_COUNT && s_stats[threadIdx.x]) atomicAdd(&p.stats[threadIdx.x], s_stats[threadIdx.x]);
}

// FootsiesEnv.reset / RESET command for the envs selected by mask (NULL = all).
template <bool P1BOT, bool P2BOT>
__global__ void __launch_bounds__(kThreads) reset_kernel(const Params p) {
    __shared__ Tables T;
    load_tables(&T, p.tables);
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
__global__ void __launch_bounds__(kThreads) seed_kernel(const Params p) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += gridDim.x * blockDim.x) {
        if (p.mask && !p.mask[i]) continue;
        uint32_t s0 = (uint32_t)(int32_t)(p.seed_base + p.first_env_index + i);
        uint32_t s1 = s0 * 1812433253u + 1u, s2 = s1 * 1812433253u + 1u, s3 = s2 * 1812433253u + 1u;
        p.pl_rng[i] = make_uint4(s0, s1, s2, s3);
    }
}

// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
thread_local char g_err[512] = "";
int fail(int code, const char *fmt, const char *detail = "") {
    snprintf(g_err, sizeof g_err, fmt, detail);
    return code;
}
#define CUDA_TRY(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) \
    return fail(FG_ERR_CUDA, #expr ": %s", cudaGetErrorString(_e)); } while (0)

void build_tables(Tables &t) {
    memset(&t, 0, sizeof t);
    static const uint32_t rows[FT_NUM_ROWS][4] = FT_ROWS_INIT;
    static const uint32_t hit[8][4] = FT_HIT_INIT;
    static const uint32_t hurt[FT_NUM_HURT][2] = FT_HURT_INIT;
    static const uint32_t push[FT_NUM_PUSH][2] = FT_PUSH_INIT;
    static const uint32_t info[FT_NUM_ACTIONS] = FT_ACTION_INFO_INIT;
    static const uint32_t attack[5] = FT_ATTACK_INIT;
    static const uint8_t cum_next[FT_NUM_CUM][4] = FT_CUM_NEXT_INIT;
    static const double step_reward[4] = FT_STEP_REWARD_INIT;
    static const double term[FT_NUM_CUM][4][2] = FT_TERM_REWARD_INIT;
    for (int i = 0; i < FT_NUM_ROWS; i++) t.rows[i] = make_uint4(rows[i][0], rows[i][1], rows[i][2], rows[i][3]);
    for (int i = 0; i < 8; i++) t.hit[i] = make_uint4(hit[i][0], hit[i][1], hit[i][2], hit[i][3]);
    for (int i = 0; i < FT_NUM_HURT; i++) t.hurt[i] = make_uint2(hurt[i][0], hurt[i][1]);
    for (int i = 0; i < FT_NUM_PUSH; i++) t.push[i] = make_uint2(push[i][0], push[i][1]);
    for (int i = 0; i < FT_NUM_ACTIONS; i++) t.action_info[i] = info[i];
    for (int i = 0; i < 5; i++) t.attack[i] = attack[i];
    memcpy(t.term_reward, term, sizeof term);
    memcpy(t.step_reward, step_reward, sizeof step_reward);
    for (int i = 0; i < FT_NUM_CUM; i++) for (int k = 0; k < 4; k++) t.cum_next[i][k] = cum_next[i][k];
    // ---- BattleAI input sequences (BattleAI.cs:192-342): F = forward, B = backward, N = none ----
    enum { N = 0, F = 1, B = 2 };
    std::vector<uint8_t> mp;
    auto rep = [&](std::vector<uint8_t> &v, int val, int n) { for (int i = 0; i < n; i++) v.push_back((uint8_t)val); };
    auto dash = [&](std::vector<uint8_t> &v) { v.push_back(F); v.push_back(N); v.push_back(F); };  // :330-342 (both dashes tap FORWARD)
    int id = 1;
    uint16_t move_off[8] = {0}, move_len[8] = {0}, att_off[8] = {0}, att_len[8] = {0};
    auto begin = [&](std::vector<uint8_t> &v, uint16_t *off) { off[id] = (uint16_t)v.size(); };
    auto end = [&](std::vector<uint8_t> &v, uint16_t *off, uint16_t *len) { len[id] = (uint16_t)(v.size() - off[id]); id++; };
    begin(mp, move_off); rep(mp, N, 30); end(mp, move_off, move_len);                                   // 1 AddNeutralMovement
    begin(mp, move_off); rep(mp, F, 40); rep(mp, B, 10); rep(mp, F, 30); rep(mp, B, 10); end(mp, move_off, move_len); // 2 FarApproach1
    begin(mp, move_off); dash(mp); rep(mp, B, 25); dash(mp); rep(mp, B, 25); end(mp, move_off, move_len);             // 3 FarApproach2
    begin(mp, move_off); rep(mp, F, 30); rep(mp, B, 10); rep(mp, F, 20); rep(mp, B, 10); end(mp, move_off, move_len); // 4 MidApproach1
    begin(mp, move_off); dash(mp); rep(mp, B, 30); end(mp, move_off, move_len);                         // 5 MidApproach2
    begin(mp, move_off); rep(mp, B, 60); end(mp, move_off, move_len);                                   // 6 FallBack1
    begin(mp, move_off); dash(mp); rep(mp, B, 60); end(mp, move_off, move_len);                         // 7 FallBack2
    memcpy(t.move_pat, mp.data(), mp.size());
    std::vector<uint8_t> apv;
    const int A = 4;
    id = 1;
    begin(apv, att_off); rep(apv, 0, 30); end(apv, att_off, att_len);                                   // 1 AddNoAttack
    begin(apv, att_off); rep(apv, A, 1); rep(apv, 0, 18); end(apv, att_off, att_len);                   // 2 OneHitImmediate
    begin(apv, att_off); rep(apv, A, 1); rep(apv, 0, 3); rep(apv, A, 1); rep(apv, 0, 18); end(apv, att_off, att_len); // 3 TwoHitImmediate
    begin(apv, att_off); rep(apv, A, 60); rep(apv, 0, 1); end(apv, att_off, att_len);                   // 4 ImmediateSpecial
    begin(apv, att_off); rep(apv, A, 120); rep(apv, 0, 1); end(apv, att_off, att_len);                  // 5 DelaySpecial
    memcpy(t.att_pat, apv.data(), apv.size());
    for (int i = 1; i < 8; i++) t.mod_magic[i] = i == 1 ? 0ull : (~0ull) / (unsigned long long)i + 1ull;
    for (int i = 0; i < 8; i++) {
        t.move_meta[i] = move_off[i] | (uint32_t)move_len[i] << 16;
        t.att_meta[i] = att_off[i] | (uint32_t)att_len[i] << 16;
    }
}

}  // namespace

struct fg_handle {
    fg_config cfg;
    fg_buffers buf;
    bool bound;
    Tables *d_tables;
    int sm_count;
    int64_t launches;
    uint8_t *d_mask;       // staging for fg_reset_host
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
    p.info_frame = h->buf.info_frame; p.info_misc = (uchar4 *)h->buf.info_misc;
    p.step_mask = h->buf.step_mask;
    p.tables = h->d_tables;
    p.first_env_index = h->cfg.first_env_index;
    p.n = h->cfg.num_envs; p.frame_skip = h->cfg.frame_skip; p.autoreset = h->cfg.autoreset;
    p.stale_intro = h->cfg.stale_intro_input;
    return p;
}

template <bool KF, bool B1, bool B2>
void launch_step_d(bool dense, int grid, cudaStream_t s, const Params &p) {
    if (dense) step_kernel<KF, B1, B2, true><<<grid, kThreads, 0, s>>>(p);
    else step_kernel<KF, B1, B2, false><<<grid, kThreads, 0, s>>>(p);
}
template <bool KF>
void launch_step_k(const fg_config &c, int grid, cudaStream_t s, const Params &p) {
    const bool d = c.dense_reward != 0;
    if (c.p1_bot && c.p2_bot) launch_step_d<KF, true, true>(d, grid, s, p);
    else if (c.p1_bot) launch_step_d<KF, true, false>(d, grid, s, p);
    else if (c.p2_bot) launch_step_d<KF, false, true>(d, grid, s, p);
    else launch_step_d<KF, false, false>(d, grid, s, p);
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
    if (h->cfg.p1_bot && h->cfg.p2_bot) reset_kernel<true, true><<<grid, kThreads, 0, s>>>(p);
    else if (h->cfg.p1_bot) reset_kernel<true, false><<<grid, kThreads, 0, s>>>(p);
    else if (h->cfg.p2_bot) reset_kernel<false, true><<<grid, kThreads, 0, s>>>(p);
    else reset_kernel<false, false><<<grid, kThreads, 0, s>>>(p);
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    return FG_OK;
}

int32_t fg_step(fg_handle *h, void *stream) {
    if (int rc = check_bound(h)) return rc;
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    const Params p = make_params(h);
    const int grid = grid_for(h, FG_BLOCKS_PER_SM);
    if (h->cfg.frame_skip == 1) launch_step_k<false>(h->cfg, grid, (cudaStream_t)stream, p);
    else launch_step_k<true>(h->cfg, grid, (cudaStream_t)stream, p);
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
