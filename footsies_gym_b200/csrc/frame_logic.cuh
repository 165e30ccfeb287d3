// frame_logic.cuh -- one FOOTSIES battle frame for one env, on registers.
//
// Everything between "state planes loaded" and "state planes + outputs stored":
//   BattleCore.FixedUpdate/UpdateFightState (BattleCore.cs:201-220, 347-364)
//   -> Fighter.UpdateInput / IncrementActionFrame / UpdateActionRequest / UpdateMovement / UpdateBoxes
//      (Fighter.cs:140-324, 472-510, 546-635, 671-719)
//   -> push / wall clamp / hitbox-hurtbox collision + damage (BattleCore.cs:483-591, Fighter.cs:352-454)
//   -> in-game bot (BattleAI.cs:41-403, queried as TrainingManager.cs:59-77 does)
//   -> observation, info, reward, termination (footsies.py:336-405, 518-570)
//
// The step kernel is bound by the integer (ALU) pipe before anything else (profiles/r01_summary.md: every ALU-pipe
// instruction per env-frame cost ~0.26 us per 4 Mi-env launch; the fused-K kernels run that pipe at ~80 %), so this
// file is written to minimise LOP3 / SHF / ISETP / SEL counts: fields are tested and updated in place in the packed
// words, table byte offsets are pre-positioned inside the row words, input processing (Attack run length, dash
// detection) and -- in the fused-K kernels -- the request decision are shared-memory table steps, multiplications by
// powers of two and adds of disjoint bit-fields go to the FMA pipe (IMAD), and facts about (action, frame) are looked
// up once per fighter per frame.
//
// fp32 discipline: compiled with -fmad=false; every add/mul rounds on its own exactly like the scalar C# expression
// it restates.  Multiplications by the facing sign (+-1) and by 0.5 are exact.
//
// The same source compiles for the host (tests/host_emulation, TEST INFRASTRUCTURE ONLY: lets the parity suite run
// this logic against the oracle on machines without a GPU; the product library never contains a host path).
#ifndef FOOTSIES_B200_FRAME_LOGIC_CUH
#define FOOTSIES_B200_FRAME_LOGIC_CUH

#include <stdint.h>

#include "state_codec.h"

#if defined(__CUDACC__)
#define FG_DEV __device__ __forceinline__
#define FG_ALIGN16 __align__(16)
namespace fg {
FG_DEV float u2f(uint32_t u) { return __uint_as_float(u); }
FG_DEV uint32_t f2u(float f) { return __float_as_uint(f); }
FG_DEV uint32_t umulhi64(uint32_t r, unsigned long long m) { return (uint32_t)__umul64hi((unsigned long long)r, m); }
FG_DEV uint32_t byte_lut(uint32_t lo, uint32_t hi, uint32_t idx) { return __byte_perm(lo, hi, idx); }
FG_DEV uint32_t umin(uint32_t a, uint32_t b) { return min(a, b); }
}  // namespace fg
#else
#include <math.h>
#include <string.h>
#define FG_DEV static inline
#define FG_ALIGN16 alignas(16)
struct uint2 { uint32_t x, y; };
struct FG_ALIGN16 uint4 { uint32_t x, y, z, w; };
namespace fg {
FG_DEV float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
FG_DEV uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
FG_DEV uint32_t umulhi64(uint32_t r, unsigned long long m) { return (uint32_t)(((unsigned __int128)r * m) >> 64); }
FG_DEV uint32_t byte_lut(uint32_t lo, uint32_t hi, uint32_t idx) {     // PRMT with selector nibbles < 8
    const unsigned long long t = (unsigned long long)hi << 32 | lo;
    uint32_t r = 0;
    for (int k = 0; k < 4; k++) r |= (uint32_t)((t >> (8 * ((idx >> (4 * k)) & 7u))) & 0xffu) << (8 * k);
    return r;
}
FG_DEV uint32_t umin(uint32_t a, uint32_t b) { return a < b ? a : b; }
}  // namespace fg
#endif

namespace fg {

// action indices (moves.py order)
enum : uint32_t { STAND = FT_IDX_STAND, FORWARD = FT_IDX_FORWARD, BACKWARD = FT_IDX_BACKWARD,
                  DASH_FORWARD = FT_IDX_DASH_FORWARD, DASH_BACKWARD = FT_IDX_DASH_BACKWARD,
                  N_ATTACK = FT_IDX_N_ATTACK, B_ATTACK = FT_IDX_B_ATTACK, N_SPECIAL = FT_IDX_N_SPECIAL,
                  B_SPECIAL = FT_IDX_B_SPECIAL, DAMAGE = FT_IDX_DAMAGE, GUARD_BREAK = FT_IDX_GUARD_BREAK,
                  GUARD_PROXIMITY = FT_IDX_GUARD_PROXIMITY, DEAD = FT_IDX_DEAD, WIN = FT_IDX_WIN };

// ---- bot input patterns (BattleAI.cs:192-342) ----
// move pattern ids: 1 Neutral, 2 FarApproach1, 3 FarApproach2, 4 MidApproach1, 5 MidApproach2, 6 FallBack1, 7 FallBack2
// attack pattern ids: 1 NoAttack, 2 OneHitImmediate, 3 TwoHitImmediate, 4 ImmediateSpecial, 5 DelaySpecial
constexpr int kMovePatBytes = 408;
constexpr int kAttPatBytes = 256;

struct BoxCfg { uint32_t h0cx, h0hw, h1cx, h1hw, pcx, pw, r0, r1; };          // 32 B (fp32 bit patterns)
struct AttackRow { uint32_t pcx, phw, rcx, rhw, pybits, rybits, result, r0; };  // 32 B
// SelectMovement / SelectAttack decision tables per distance bucket (BattleAI.cs:68-190)
struct BotRow { unsigned long long magic_m, magic_a; uint32_t n_m, sel_m, n_a, sel_a; };   // 32 B

struct FG_ALIGN16 Tables {
    uint4 rows[FT_NUM_ROWS];              // [action * 64 + frame], see tools/gen_kernel_tables.py
    BoxCfg boxcfg[16];
    AttackRow attack[8];
    BotRow bot[8];                        // [distance bucket]
    double term_reward[FT_NUM_CUM][4][2];
    double step_reward[4];
    uint8_t cum_next[16][4];
    uint32_t move_meta[8], att_meta[8];   // pattern offset | length << 16
    uint8_t bucket_of[16];                // [clamp(ceil(2 * distance), 4, 9) - 4] -> bucket
    uint16_t dash_fsm[256][4];            // [state][Left | Right << 1] -> next state | dash-by-Left << 8 | dash-by-Right << 9
    uint8_t arun_lut[64 * 8];             // [run * 8 + input] -> new run | special << 6 | attack-down << 7
    uint8_t req_lut[2][1024];             // [side][inputs + flags] -> requested action | FREE | WANT_BUFFER | ENDED
    uint8_t move_pat[2][kMovePatBytes];   // [side] InputDefine bits (P1: forward = Right, P2: forward = Left)
    uint8_t att_pat[kAttPatBytes];
};
static_assert(sizeof(Tables) % 16 == 0, "Tables is copied as uint4");
static_assert(sizeof(BoxCfg) == 32 && sizeof(AttackRow) == 32 && sizeof(BotRow) == 32, "byte offsets are pre-positioned");

// Per-thread packed statistics: 32-bit words of four 8-bit lanes each, bumped once per frame and folded by a warp
// reduction when a thread has gone kStatFlushFrames frames without a flush (and at kernel end).
//   a: episodes | P1 wins | P2 wins | double KOs      (1 << 8*w trick: w = winner code)
//   r: (none)   | hits    | blocks  | guard breaks    (1 << 8*DamageResult for each of the two attack passes)
//   s: specials | specials-from-neutral | resets | env-frames simulated (bumped by the step loop)
//   ep_frames: sum of the lengths of the episodes that ended (plain 32-bit sum)
struct StatAcc { uint32_t a, r, s, ep_frames; };
constexpr uint32_t kStatFlushFrames = 120u;   // <= 2 events per lane per frame -> a byte lane cannot overflow

struct Env {            // one battle, in registers
    float pos1, vel1, pos2, vel2;
    uint32_t pk1, hist1, pk2, hist2;
    int32_t frame;
    uint32_t misc, bq2, bq1;
    uint32_t r0, r1, r2, r3;
    bool drew;          // the RNG advanced since the state was loaded (the RNG plane is only written back then)
};

// packed-word masks
constexpr uint32_t M_FRAME1 = 1u << FGP_FRAME_SHIFT, M_ACT = 31u << FGP_ACT_SHIFT, M_STUN = 31u << FGP_STUN_SHIFT,
                   M_STUN1 = 1u << FGP_STUN_SHIFT, M_GUARD = 3u << FGP_GUARD_SHIFT, M_GUARD1 = 1u << FGP_GUARD_SHIFT,
                   M_VITAL = 1u << FGP_VITAL_SHIFT, M_HIT = 1u << FGP_HITCNT_SHIFT, M_BUF = 1u << FGP_BUF_SHIFT,
                   M_RSV = 1u << FGP_RSV_SHIFT, M_INBACK = 1u << FGP_INBACK_SHIFT, M_RPROX = 1u << FGP_RPROX_SHIFT,
                   M_SHMAG = 7u << FGP_SHAKE_SHIFT, M_SHMAG1 = 1u << FGP_SHAKE_SHIFT, M_SHSIGN = 1u << FGP_SHAKE_SIGN_SHIFT;

// UnityEngine.Random restated as xorshift128 (closed source; see DESIGN.md "parity unpinned")
FG_DEV uint32_t rng_next(Env &e) {
    uint32_t t = e.r0 ^ (e.r0 << 11);
    e.r0 = e.r1; e.r1 = e.r2; e.r2 = e.r3;
    e.r3 = e.r3 ^ (e.r3 >> 19) ^ t ^ (t >> 8);
    e.drew = true;
    return e.r3;
}

struct FrameOut {       // per-fighter products of the pre-collision phases
    uint32_t z, w;      // row words of the (action, frame) the boxes are built from
    float pos_b;        // position when the boxes were built (after movement, before push)
    bool special_started;   // this frame's request switched the fighter into N_SPECIAL / B_SPECIAL from another action
    bool from_neutral;      // ... and the previous action was not N_ATTACK / B_ATTACK
};

// Fighter.UpdateInput + IncrementActionFrame + UpdateActionRequest + UpdateMovement for one fighter.
// SIDE 0 = P1 (faces right: forward = Right), 1 = P2 (faces left: forward = Left).  `hist` is the fighter's input word:
// dash-automaton state [0:8) | Attack run length [8:14).
// LUT_REQUEST selects how the request is decided: by one more table lookup (fewest ALU-pipe instructions: the fused-K
// kernels, which are purely ALU-bound, gain 4 %) or by a short select chain on predicates (one dependent shared-memory
// access less per fighter: the K = 1 kernels, which run 24 warps per SM against HBM latency, are 3 % faster with it).
template <int SIDE, bool LUT_REQUEST>
FG_DEV void update_fighter(const Tables &T, uint32_t in, float &pos, float &vel, uint32_t &pk, uint32_t &hist, FrameOut &fo) {
    // ---- UpdateInput (Fighter.cs:172-188) + CheckSpecialAttackInput (:569-583) + CheckForward/BackwardDashInput
    //      (:585-635), each as one table step: the Attack run length (new run | special | attack-down) and the
    //      dash-detection automaton over the Left/Right bits (next state | dash-by-Left | dash-by-Right) ----
    const uint32_t in_lr = in & 3u;
    const uint32_t au = T.arun_lut[((hist >> (FGH_ARUN_SHIFT - 3)) & (63u << 3)) + in];
    const uint32_t du = T.dash_fsm[hist & 255u][in_lr];
    hist = (du & 255u) | (au & 63u) << FGH_ARUN_SHIFT;
    // ---- IncrementActionFrame (Fighter.cs:140-166): sprite shake decays (sign flips, magnitude - 1); hit stun ticks
    //      down and freezes the frame counter, else the frame counter advances ----
    uint32_t delta = (pk & M_STUN) ? (0u - M_STUN1) : M_FRAME1;
    if (pk & M_SHMAG) { delta -= M_SHMAG1; pk ^= M_SHSIGN; }
    pk += delta;
    const bool stun0 = (pk & M_STUN) == 0u;

    // ---- UpdateActionRequest (Fighter.cs:201-286) with the RequestAction chain (Fighter.cs:472-510) collapsed: when the
    //      action ended or is alwaysCancelable ("free") the FIRST request of the chain wins, otherwise the only effect
    //      a request can have is buffering N_SPECIAL inside a cancel window.  The carry bit END is only ever set while
    //      the fighter is out of hit stun, so it means "currentActionFrame >= frameCount now" (Fighter.cs:90). ----
    uint32_t req;
    bool differs, set, want_buffer;
    if (LUT_REQUEST) {
        // one lookup (tools/gen_kernel_tables.py build_request_lut); its index is assembled from bit-fields that are
        // already integers: this frame's Left/Right, the run-length and dash-automaton outputs, and the packed word's
        // isReserveProximityGuard and carry bits
        static_assert(FGP_RPROX_SHIFT == 23 && FGP_CARRY_END == (1u << 28), "request LUT index assembly");
        const uint32_t ridx = in_lr | ((pk >> 21) & 0x384u) | ((au >> 3) & 0x18u) | ((du >> 3) & 0x60u);
        const uint32_t rq = T.req_lut[SIDE][ridx];
        req = rq & 15u;
        differs = ((pk ^ (req << FGP_ACT_SHIFT)) & M_ACT) != 0u;
        set = (rq & FT_REQ_FREE) && ((rq & FT_REQ_ENDED) || differs);
        want_buffer = (rq & FT_REQ_WANT_BUFFER) != 0u;
    } else {
        const bool special = (au & 64u) != 0u;                          // Attack released after >= 59 held frames
        const bool atk_down = (au & 128u) != 0u;                        // IsAttackInput(inputDown[0])
        const bool dash_f = (du & (SIDE == 0 ? 0x200u : 0x100u)) != 0u; // P1's forward is Right, P2's is Left
        const bool dash_b = (du & (SIDE == 0 ? 0x100u : 0x200u)) != 0u;
        const bool ended = (pk & FGP_CARRY_END) != 0u;
        const bool in_normal = (pk & FGP_CARRY_NORMAL) && !ended;
        const uint32_t dir = in_lr != 0u ? 1u : 0u;
        // attack request: N_ATTACK 5 / B_ATTACK 6 / N_SPECIAL 7 / B_SPECIAL 8 (the B_ variant when a direction is held)
        const uint32_t areq = special ? N_SPECIAL + dir : in_normal ? (uint32_t)N_SPECIAL : N_ATTACK + dir;
        // movement request by (Left, Right, isReserveProximityGuard): byte LUT in two registers (PRMT)
        const uint32_t mv = byte_lut(SIDE == 0 ? 0x00010200u : 0x00020100u, SIDE == 0 ? 0x00010e00u : 0x000e0100u,
                                     in_lr | ((pk >> (FGP_RPROX_SHIFT - 2)) & 4u));
        req = (special || atk_down) ? areq : dash_f ? (uint32_t)DASH_FORWARD : dash_b ? (uint32_t)DASH_BACKWARD : mv;
        const bool free_to_switch = ended || (pk & FGP_CARRY_ALWAYS);
        differs = ((pk ^ (req << FGP_ACT_SHIFT)) & M_ACT) != 0u;
        set = free_to_switch && (ended || differs);
        want_buffer = !free_to_switch && req == N_SPECIAL;
    }
    const bool normal = (pk & FGP_CARRY_NORMAL) != 0u;                  // current action is N_ATTACK / B_ATTACK
    // Forced requests take precedence, both only once hit stun is over: the reserved GUARD_BREAK (Fighter.cs:212-218),
    // else the buffered cancel into N_SPECIAL after a connected hit (Fighter.cs:222-229); they return before the
    // isInputBackward / isReserveProximityGuard bookkeeping.
    if ((pk & (M_RSV | M_BUF)) && stun0 && ((pk & M_RSV) || (pk & M_HIT))) {
        req = (pk & M_RSV) ? (uint32_t)GUARD_BREAK : (uint32_t)N_SPECIAL;
        set = true; differs = true; want_buffer = false;
    } else {
        // isInputBackward = holding back now; the proximity-guard reservation is consumed (Fighter.cs:271-285)
        pk = (pk & ~(M_INBACK | M_RPROX)) | (in & (SIDE == 0 ? 1u : 2u)) << (FGP_INBACK_SHIFT - (SIDE == 0 ? 0 : 1));
    }
    if (SIDE == 0) {                                                    // statistics are about P1 only
        fo.special_started = set && differs && (req - N_SPECIAL) < 2u;
        fo.from_neutral = !normal;
    }
    if (set) pk = (pk & (M_STUN | M_GUARD | M_VITAL | M_INBACK | M_RPROX)) + (req << FGP_ACT_SHIFT);   // SetCurrentAction (Fighter.cs:546-563)

    // ---- frame data of the (action, frame) the fighter ends up in: the low bits of the packed word are the row index ----
    const uint4 row = T.rows[pk & FGP_ROW_MASK];
    if (want_buffer) pk |= (row.z & FT_Z_CANCEL) << (FGP_BUF_SHIFT - 3);   // cancel window (Fighter.cs:492-505)
    // refresh the carry bits; END only while out of hit stun (a stunned fighter's frame counter does not advance)
    pk = (pk & ~FGP_CARRY_MASK) | (row.w & (stun0 ? FGP_CARRY_MASK : (FGP_CARRY_MASK & ~FGP_CARRY_END)));

    // ---- UpdateMovement (Fighter.cs:291-319) ----
    if (stun0) {
        const float dx = u2f(row.x);
        pos = pos + (SIDE == 0 ? dx : -dx);
        if (row.z & FT_Z_HAS_MOVEMENT) vel = u2f(row.y);
    }
    fo.z = row.z;
    fo.w = row.w;
    fo.pos_b = pos;
}

// Table rows addressed by byte offsets that sit pre-positioned in the row words.
FG_DEV const BoxCfg &boxcfg_of(const Tables &T, uint32_t z) {
    return *reinterpret_cast<const BoxCfg *>(reinterpret_cast<const char *>(T.boxcfg) + (z & (15u << FT_Z_BOXCFG_SHIFT)));
}
FG_DEV const AttackRow &attack_of(const Tables &T, uint32_t w) {
    return *reinterpret_cast<const AttackRow *>(reinterpret_cast<const char *>(T.attack) + (w & (7u << FT_W_KIND_SHIFT)));
}
static_assert(FT_Z_BOXCFG_SHIFT == 5 && FT_W_KIND_SHIFT == 5, "id << 5 == byte offset of a 32-byte table row");
static_assert(FT_Z_CANCEL == 8u, "update_fighter shifts the cancel bit into the buffered-cancel bit");

// World x-extent of a box built at position pos_b (Fighter.cs:706-719: x = pos + data.x * sign; BoxBase xMin/xMax,
// Fighter.cs:12-13) and then displaced by the push (s) and the wall clamp (t) like ApplyPositionChange does to
// already-built boxes (Fighter.cs:331-350).  Every operation rounds separately.
template <int SIDE>
FG_DEV void box_extent(float pos_b, uint32_t cx_bits, uint32_t hw_bits, float s, float t, float &lo, float &hi) {
    const float cx = u2f(cx_bits), hw = u2f(hw_bits);
    const float x = ((pos_b + (SIDE == 0 ? cx : -cx)) + s) + t;
    lo = x - hw;
    hi = x + hw;
}

// Geometry half of BattleCore.UpdateHitboxHurtboxCollision (BattleCore.cs:535-565) for one attacker: does its real /
// proximity hitbox overlap any of the victim's (<= 2) hurtboxes?  BoxBase.Overlaps (Fighter.cs:17-25, inclusive); the
// y half of each test is pre-resolved into the victim row's y-bits over hit boxes.  Straight-line code: the boxes
// are a snapshot, so both attackers' tests can be evaluated before either attack is applied.
template <int ASIDE>
FG_DEV void attack_overlaps(const AttackRow &ar, const BoxCfg &vb, const FrameOut &af, const FrameOut &vf, float a_s,
                            float a_t, float v_s, float v_t, bool &real_hit, bool &prox_hit) {
    float plo, phi, rlo, rhi, v0lo, v0hi, v1lo, v1hi;
    box_extent<ASIDE>(af.pos_b, ar.pcx, ar.phw, a_s, a_t, plo, phi);
    box_extent<ASIDE>(af.pos_b, ar.rcx, ar.rhw, a_s, a_t, rlo, rhi);
    box_extent<1 - ASIDE>(vf.pos_b, vb.h0cx, vb.h0hw, v_s, v_t, v0lo, v0hi);
    box_extent<1 - ASIDE>(vf.pos_b, vb.h1cx, vb.h1hw, v_s, v_t, v1lo, v1hi);
    constexpr uint32_t Y0 = 0xffu << FT_Z_YMASK0_SHIFT, Y1 = 0xffu << FT_Z_YMASK1_SHIFT;
    // otherBox.xMax >= xMin && otherBox.xMin <= xMax with self = hitbox, other = hurtbox; an absent box has no y-bits
    const bool r0 = (vf.z & ar.rybits & Y0) && v0hi >= rlo && v0lo <= rhi;
    const bool r1 = (vf.z & ar.rybits & Y1) && v1hi >= rlo && v1lo <= rhi;
    const bool p0 = (vf.z & ar.pybits & Y0) && v0hi >= plo && v0lo <= phi;
    const bool p1 = (vf.z & ar.pybits & Y1) && v1hi >= plo && v1lo <= phi;
    real_hit = (af.z & FT_Z_REAL) && (r0 || r1);
    prox_hit = (af.z & FT_Z_PROX) && (p0 || p1);
}

// Effect half of one attacker -> victim pass (BattleCore.cs:567-586): NotifyAttackHit / NotifyDamaged /
// GetHitStunFrame / SetHitStun / SetSpriteShakeFrame / NotifyInProximityGuardRange (Fighter.cs:352-454).
// Hit counts are the CURRENT ones (P1's hit may just have changed P2), boxes and the victim's blocking stance are
// the snapshot (an attacker's own action is never changed by its attack).
// Returns the DamageResult (0 none, 1 damage, 2 guard, 3 guard break) | 4 when the victim's guard bar dropped.
template <int ASIDE>
FG_DEV uint32_t attack_apply(uint32_t result, uint32_t &apk, uint32_t &vpk, const FrameOut &af, const FrameOut &vf,
                             bool real_hit, bool prox_hit, StatAcc &acc) {
    const bool can = (af.z & (FT_Z_PROX | FT_Z_REAL)) && !(apk & M_HIT);   // a hitbox is out and CanAttackHit
    if (can && real_hit) {
        const bool brk = (vpk & M_GUARD) == 0u;                         // guardHealth < 0 after the decrement
        const bool guarding = (vf.z & FT_Z_GUARDING) != 0u;             // BACKWARD or a Type == Guard action
        uint32_t nact, stun, res;
        uint32_t keep = vpk & (M_GUARD | M_VITAL | M_INBACK | M_RPROX);
        if (!brk) keep -= M_GUARD1;
        if (guarding) {
            nact = (result >> 5) & 31u;
            stun = brk ? (result >> 21) & 31u : (result >> 16) & 31u;
            if (brk) keep |= M_RSV;
            res = brk ? 3u : 6u;                                        // 2 | guard dropped
            acc.r += brk ? (1u << 24) : (1u << 16);                     // byte lane = DamageResult: 3 guard break, 2 guard
        } else {
            if (result & (1u << 10)) keep &= ~M_VITAL;
            nact = result & 31u;
            stun = (result >> 11) & 31u;
            res = brk ? 1u : 5u;
            acc.r += 1u << 8;                                           // 1 damage
        }
        // SetSpriteShakeFrame: min(stun / 3, 6), the victim of P1 faces left -> positive, of P2 -> negative
        const uint32_t mag = umin(stun / 3u, 6u);
        vpk = keep | nact << FGP_ACT_SHIFT | stun << FGP_STUN_SHIFT | mag << FGP_SHAKE_SHIFT
            | (ASIDE == 0 ? 0u : (mag ? M_SHSIGN : 0u));
        // the victim's new action starts at frame 0: never END (every hit action lasts >= 15 frames), never
        // ALWAYS / NORMAL -> carry bits 0
        apk = (apk & ~(M_STUN | (stun ? FGP_CARRY_END : 0u))) | stun << FGP_STUN_SHIFT | M_HIT;   // END is void in hit stun
        return res;
    }
    // NotifyInProximityGuardRange: latch only while the victim holds back (Fighter.cs:400-406)
    if (can && prox_hit) vpk |= (vpk & M_INBACK) << (FGP_RPROX_SHIFT - FGP_INBACK_SHIFT);
    return 0u;
}

// Distance bucket index of bot_next (0..5): ceil(clamp(2 * dist, 4, 9)) - 4; bot_dist_of_index is a distance with that index.
FG_DEV uint32_t bot_dist_index(float dist) {
    float t2 = dist * 2.0f;
    t2 = t2 < 4.0f ? 4.0f : t2 > 9.0f ? 9.0f : t2;
    return (uint32_t)((int)ceilf(t2) - 4);
}
FG_DEV float bot_dist_of_index(uint32_t idx) { return 0.5f * (float)(idx + 4u); }

// BattleAI.getNextAIInput (BattleAI.cs:41-66) on pattern-position + remaining-count queues.  `dist` and `opp_act`
// are the state captured by the PREVIOUS call (the ascending shift loop at BattleAI.cs:358-361 makes fightStates[5]
// exactly that).  r % n for 2 <= n <= 7 without a division: floor(r / n) == umul64hi(r, floor(2^64 / n) + 1).
template <int SIDE>
FG_DEV uint32_t bot_next(const Tables &T, Env &e, uint32_t &q, float dist, uint32_t opp_act) {
    // queue word: move position in move_pat [0:9) | moves remaining [9:16) | attack position in att_pat [16:24) |
    // attacks remaining [24:31): dequeuing is one add on the packed word
    uint32_t input = 0u;
    const bool have_m = (q & (127u << 9)) != 0u, have_a = (q & (127u << 24)) != 0u;
    if (have_m) {
        input = T.move_pat[SIDE][q & 511u];
        q += 1u - (1u << 9);
    }
    if (have_a) {
        input |= T.att_pat[(q >> 16) & 255u];
        q += (1u << 16) - (1u << 24);
    }
    if (!(have_m && have_a)) {                                          // an empty queue is refilled and contributes 0 (BattleAI.cs:50-62)
        // distance buckets > 4, > 3, > 2.5, > 2, else (BattleAI.cs:70-124): 2 * dist is exact, ceil() turns the
        // strict comparisons into a table index
        const BotRow &b = T.bot[T.bucket_of[bot_dist_index(dist)]];
        if (!have_m) {                                                  // SelectMovement (BattleAI.cs:68-126)
            const uint32_t r = rng_next(e);
            const uint32_t k = r - umulhi64(r, b.magic_m) * b.n_m;      // Random.Range(0, n)
            const uint32_t meta = T.move_meta[(b.sel_m >> (4u * k)) & 15u];   // offset | length << 16
            q = (q & 0xffff0000u) | (meta & 0xffffu) | (meta >> 16) << 9;
        }
        if (!have_a) {                                                  // SelectAttack (BattleAI.cs:128-190)
            const bool opp_hurt = opp_act == DAMAGE || opp_act == GUARD_BREAK || opp_act == N_SPECIAL || opp_act == B_SPECIAL;
            const bool opp_normal = opp_act == N_ATTACK || opp_act == B_ATTACK;
            uint32_t ap = 3u;                                           // AddTwoHitImmediateAttack, no draw
            if (!(opp_hurt || (opp_normal && (b.n_a >> 8)))) {          // n_a bit 8: the "mid" bucket that punishes normals
                const uint32_t r = rng_next(e);
                const uint32_t n = b.n_a & 255u;
                const uint32_t k = r - umulhi64(r, b.magic_a) * n;
                ap = (b.sel_a >> (4u * k)) & 15u;
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
// P1's bot (by_example) survives too: the game is launched with --p1-bot --p1-spectator (footsies.py:230-232), the bot
// actor is wrapped in a TrainingActorRemoteSpectator (GameManager.cs:200-201), and `actorP1 is TrainingBattleAIActor`
// (BattleCore.cs:274) is false for the wrapper -- BattleAI.Reset() is only ever called for P2's bot.  P1's bot therefore
// keeps its queues, decides its first input of the new round on the FightState it recorded last (before the terminal
// frame, or the current state on a RESET in mid-round), and its very first query (fightStates still null,
// BattleAI.cs:30,47) returns 0 without drawing.  Found and pinned by tests/test_oracle_vs_ref.py.
template <bool P1BOT, bool P2BOT>
FG_DEV void reset_env(const Tables &T, Env &e, bool stale_intro) {
    const bool was_done = (e.misc >> FGM_DONE_SHIFT) & 1u;
    uint32_t a1 = (e.misc >> FGM_ACTOR1_SHIFT) & 7u, a2 = (e.misc >> FGM_ACTOR2_SHIFT) & 7u;
    if (!stale_intro) { a1 = 0u; a2 = 0u; }
    float p1_dist = 0.0f;
    uint32_t p1_opp = 0u;
    bool p1_called = false;
    if (P1BOT) {
        p1_called = (e.misc >> FGM_P1CALLED_SHIFT) & 1u;
        p1_dist = was_done ? bot_dist_of_index((e.misc >> FGM_P1MEM_SHIFT) & 7u) : fabsf(e.pos2 - e.pos1);
        p1_opp = was_done ? (e.misc >> (FGM_P1MEM_SHIFT + 3)) & 31u : (e.pk2 >> FGP_ACT_SHIFT) & 31u;
    }
    uint32_t pk[2] = { e.pk1, e.pk2 };
    uint32_t npk[2];
#pragma unroll
    for (int s = 0; s < 2; s++) {
        uint32_t stun = (pk[s] >> FGP_STUN_SHIFT) & 31u;
        uint32_t keep = pk[s] & (M_INBACK | M_RPROX);
        if (was_done) {
            // the End-state frame (BattleCore.cs:371-381) ran once: hit stun ticks; a dead fighter went through the
            // normal request path with cleared inputs (flags reset), a winner returned early (flags kept)
            if (stun > 0u) stun--;
            if (!(pk[s] & M_VITAL)) keep = 0u;
        }
        // Intro frame: IncrementActionFrame (frame 0 -> 1 unless in hit stun), RequestAction(STAND) is a no-op
        uint32_t frame = 1u;
        if (stun > 0u) { stun--; frame = 0u; }
        npk[s] = STAND << FGP_ACT_SHIFT | frame << FGP_FRAME_SHIFT | stun << FGP_STUN_SHIFT | 3u << FGP_GUARD_SHIFT
               | M_VITAL | keep | FGP_CARRY_ALWAYS;                     // STAND: alwaysCancelable, 24 frames
    }
    e.pk1 = npk[0]; e.pk2 = npk[1];
    e.pos1 = -2.0f; e.pos2 = 2.0f; e.vel1 = 0.0f; e.vel2 = 0.0f;
    // UpdateInput(stale input) after ClearInput: one automaton / run-length step from the cleared state
    e.hist1 = (T.dash_fsm[0][a1 & 3u] & 255u) | ((a1 >> 2) & 1u) << FGH_ARUN_SHIFT;
    e.hist2 = (T.dash_fsm[0][a2 & 3u] & 255u) | ((a2 >> 2) & 1u) << FGH_ARUN_SHIFT;
    e.frame = -1;
    e.bq2 = 0u;                                                        // BattleAI.Reset (BattleAI.cs:393-403): P2's bot only
    if (!P1BOT) e.bq1 = 0u;
    // first bot query at the Fight transition (BattleCore.cs:289), P1 first: P2's decision input is the round-start state
    if (P1BOT) a1 = p1_called ? bot_next<0>(T, e, e.bq1, p1_dist, p1_opp) : 0u;
    if (P2BOT) a2 = bot_next<1>(T, e, e.bq2, 4.0f, STAND);
    e.misc = a1 << FGM_ACTOR1_SHIFT | a2 << FGM_ACTOR2_SHIFT            // recorded inputs 0, done 0, cum 0
           | (P1BOT ? 1u << FGM_P1CALLED_SHIFT : 0u);
}

// What one env-step hands back: FootsiesEnv._extract_obs / _extract_info (footsies.py:336-380) incl. the
// DEAD/WIN -> STAND remap of step() (footsies.py:538-549; a no-op on the reset observation, which is always STAND).
struct StepOutputs {
    float obs[8];       // guard p1,p2 | move index p1,p2 | move_frame p1,p2 | position p1,p2
    uint32_t info_misc; // p1_action | p2_action << 8 | p1_hitstun << 16 | p2_hitstun << 24
};
FG_DEV void make_outputs(const Env &e, StepOutputs &o) {
    uint32_t m1 = (e.pk1 >> FGP_ACT_SHIFT) & 31u, m2 = (e.pk2 >> FGP_ACT_SHIFT) & 31u;
    if (m1 >= DEAD) m1 = STAND;
    if (m2 >= DEAD) m2 = STAND;
    const uint32_t f1 = m1 <= BACKWARD ? 0u : e.pk1 & 63u;
    const uint32_t f2 = m2 <= BACKWARD ? 0u : e.pk2 & 63u;
    o.obs[0] = (float)((e.pk1 >> FGP_GUARD_SHIFT) & 3u); o.obs[1] = (float)((e.pk2 >> FGP_GUARD_SHIFT) & 3u);
    o.obs[2] = (float)m1; o.obs[3] = (float)m2;
    o.obs[4] = (float)f1; o.obs[5] = (float)f2; o.obs[6] = e.pos1; o.obs[7] = e.pos2;
    o.info_misc = ((e.misc >> FGM_REC1_SHIFT) & 7u) | ((e.misc >> FGM_REC2_SHIFT) & 7u) << 8
                | ((e.pk1 >> FGP_STUN_SHIFT) & 31u) << 16 | ((e.pk2 >> FGP_STUN_SHIFT) & 31u) << 24;
}

// FootsiesFrameSkipped._is_obs_skippable (wrappers/frame_skip.py:56-66) on the state the observation is made from: P1 is
// in the middle of a move (observed move_frame != 0, i.e. not STAND / FORWARD / BACKWARD and past frame 0) while P2 is not
// in a hit / guard move, or P1 is in DAMAGE.  Same DEAD / WIN -> STAND remap as make_outputs.
FG_DEV bool obs_is_skippable(const Env &e) {
    uint32_t m1 = (e.pk1 >> FGP_ACT_SHIFT) & 31u, m2 = (e.pk2 >> FGP_ACT_SHIFT) & 31u;
    if (m1 >= DEAD) m1 = STAND;
    if (m2 >= DEAD) m2 = STAND;
    const uint32_t f1 = m1 <= BACKWARD ? 0u : e.pk1 & 63u;
    const bool p2_hit_guard = m2 == DAMAGE || m2 == FT_IDX_GUARD_STAND || m2 == FT_IDX_GUARD_CROUCH || m2 == FT_IDX_GUARD_M || m2 == GUARD_BREAK;
    return (f1 != 0u && !p2_hit_guard) || m1 == DAMAGE;
}

// One fight frame for one env (everything between "inputs known" and "state after the frame").
// Sets `terminal`, accumulates the Python float64 reward into `reward`, bumps the packed statistics.
template <bool P1BOT, bool P2BOT, bool DENSE, bool LUT_REQUEST>
FG_DEV void simulate_frame(const Tables &T, Env &e, uint32_t in1, uint32_t in2, double &reward, bool &terminal, StatAcc &acc) {
    // state the bots will be shown after this frame (previous call's capture == state before this frame)
    float pre_dist = 0.0f;
    uint32_t pre_a1 = 0u, pre_a2 = 0u;
    if (P1BOT || P2BOT) {
        pre_dist = fabsf(e.pos2 - e.pos1);
        pre_a1 = (e.pk1 >> FGP_ACT_SHIFT) & 31u;
        pre_a2 = (e.pk2 >> FGP_ACT_SHIFT) & 31u;
    }

    e.frame++;
    // BattleCore.RecordInput (BattleCore.cs:593-607): recording stops after maxRecordingInputFrame frames
    if (e.frame < FG_MAX_RECORDING_INPUT_FRAME)
        e.misc = (e.misc & ~(63u << FGM_REC1_SHIFT)) | (in1 + in2 * 8u) << FGM_REC1_SHIFT;

    FrameOut f1, f2;
    update_fighter<0, LUT_REQUEST>(T, in1, e.pos1, e.vel1, e.pk1, e.hist1, f1);
    update_fighter<1, LUT_REQUEST>(T, in2, e.pos2, e.vel2, e.pk2, e.hist2, f2);

    // ---- UpdatePushCharacterVsCharacter (BattleCore.cs:483-501), UnityEngine.Rect semantics: x = left edge, strict ----
    const BoxCfg &b1 = boxcfg_of(T, f1.z), &b2 = boxcfg_of(T, f2.z);
    const float px1 = e.pos1 + u2f(b1.pcx), w1 = u2f(b1.pw);
    const float px2 = e.pos2 - u2f(b2.pcx), w2 = u2f(b2.pw);
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
    if ((f1.z | f2.z) & (FT_Z_PROX | FT_Z_REAL)) {                      // somebody has a hitbox out
        const AttackRow &ar1 = attack_of(T, f1.w), &ar2 = attack_of(T, f2.w);
        bool real_a, prox_a, real_b, prox_b;
        attack_overlaps<0>(ar1, b2, f1, f2, s1, t1, s2, t2, real_a, prox_a);
        attack_overlaps<1>(ar2, b1, f2, f1, s2, t2, s1, t1, real_b, prox_b);
        res_a = attack_apply<0>(ar1.result, e.pk1, e.pk2, f1, f2, real_a, prox_a, acc);   // result on P2
        res_b = attack_apply<1>(ar2.result, e.pk2, e.pk1, f2, f1, real_b, prox_b, acc);   // result on P1
    }

    // P1 started a special move and still is in it after the collision phase (wrappers/statistics.py:36-46)
    if (f1.special_started && res_b == 0u) acc.s += f1.from_neutral ? 0x101u : 1u;

    // ---- KO (BattleCore.cs:212-217), termination (footsies.py:555) ----
    terminal = ((e.pk1 & e.pk2) & M_VITAL) == 0u;

    // ---- reward (footsies.py:382-405); Python floats are doubles ----
    if (DENSE) {
        const uint32_t code = (res_b >> 2) | (res_a >> 2) << 1;          // bit 0: P1's guard bar dropped, bit 1: P2's
        if (code | (terminal ? 1u : 0u)) {
            const uint32_t cum = T.cum_next[(e.misc >> FGM_CUM_SHIFT) & 15u][code];
            e.misc = (e.misc & ~(15u << FGM_CUM_SHIFT)) | cum << FGM_CUM_SHIFT;
            reward += terminal ? T.term_reward[cum][code][(e.pk2 & M_VITAL) ? 0 : 1] : T.step_reward[code];
        }
    } else if (terminal) {
        reward += (e.pk2 & M_VITAL) ? -1.0 : 1.0;
    }

    uint32_t n1 = in1, n2 = in2;
    if (terminal) {
        // ChangeRoundState(KO): ClearInput on both fighters (BattleCore.cs:292-299); actors keep their inputs
        const bool dead1 = !(e.pk1 & M_VITAL), dead2 = !(e.pk2 & M_VITAL);
        e.hist1 = 0u; e.hist2 = 0u;
        acc.a += 1u + (1u << (8u * (dead1 && dead2 ? 3u : dead2 ? 1u : 2u)));
        acc.ep_frames += (uint32_t)(e.frame + 1);
        e.misc |= 1u << FGM_DONE_SHIFT;
        // P1's never-Reset() bot is not asked on the terminal frame; what it recorded last (= the state before this frame)
        // is what it will decide on at the start of the next round (see reset_env)
        if (P1BOT) e.misc = (e.misc & ~(255u << FGM_P1MEM_SHIFT)) | (bot_dist_index(pre_dist) | pre_a2 << 3) << FGM_P1MEM_SHIFT;
    } else {
        // ---- TrainingManager.Step (TrainingManager.cs:59-77): actors' inputs for the next frame; bots are asked
        //      after the frame, P1 first, and not on the terminal frame ----
        if (P1BOT) n1 = bot_next<0>(T, e, e.bq1, pre_dist, pre_a2);
        if (P2BOT) n2 = bot_next<1>(T, e, e.bq2, pre_dist, pre_a1);
    }
    e.misc = (e.misc & ~(63u << FGM_ACTOR1_SHIFT)) | (n1 + n2 * 8u) << FGM_ACTOR1_SHIFT;
}

}  // namespace fg
#endif
