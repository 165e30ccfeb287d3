/*
 * footsies_oracle.h -- CPU ORACLE for the FOOTSIES per-frame battle update.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it,
 * and there only as the checker / reported CPU baseline.  The product (footsies_gym_b200)
 * never links, imports or falls back to this code.
 *
 * What it is: a deliberately naive, scalar C restatement of the reference's battle logic
 * (/root/reference/Assets/Script/{BattleCore,Fighter,BattleAI,ActionData}.cs, the training
 * glue in TrainingManager.cs / Training*Actor.cs) and of the Python observation / reward /
 * termination code in footsies-gym/footsies_gym/envs/footsies.py.  It keeps the full
 * 180-entry input arrays, real FIFO queues and range-scanned frame data so that every
 * function can be read side by side with the C# it cites.
 *
 * PARITY STATUS:
 *   - PINNED to the reference's own source text: the C# engine cannot run here (no Unity/mono/dotnet, no game binary)
 *     and the reference ships no tests, so tools/cs2cpp.py transliterates Assets/Script/{Fighter,BattleAI,BattleCore,
 *     ActionData,...}.cs mechanically into C++ (oracle/_ref/, built by oracle/Makefile.ref from the sources where they lie
 *     under /root/reference, never committed) and tests/test_oracle_vs_ref.py steps this oracle and that engine side by
 *     side on every parity tape: every field of every trace after every step is byte-identical.  The known-answer tests
 *     (tests/test_oracle_kat.py) are asked of both engines.
 *   - The Python half (obs / info / reward / termination / frame_delay) is pinned by golden vectors produced by the
 *     reference's own unmodified FootsiesEnv class driven over its socket protocol (tests/golden/make_golden.py).
 *   - Still "PARITY UNPINNED", because it is third-party code that is NOT under /root/reference: UnityEngine.Random (closed
 *     source; restated from public descriptions as Marsaglia xorshift128), UnityEngine.Rect.Overlaps (documented behaviour),
 *     Time.deltaTime = 0.02 and Mono's evaluation of float expressions.  Both engines restate them identically.
 */
#ifndef FOOTSIES_ORACLE_H
#define FOOTSIES_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Expanded per-fighter state, superset of FighterState.cs:26-56 minus boxes. */
typedef struct {
    float pos_x;
    float velocity_x;
    int32_t action_id;        /* currentActionID (CommonActionID value, e.g. 110)      */
    int32_t action_frame;     /* currentActionFrame                                     */
    int32_t hitstun;          /* currentHitStunFrame                                    */
    int32_t guard;            /* guardHealth                                            */
    int32_t vital;            /* vitalHealth                                            */
    int32_t hit_count;        /* currentActionHitCount                                  */
    int32_t buffer_id;        /* bufferActionID (-1 = none)                             */
    int32_t reserve_id;       /* reserveDamageActionID (-1 = none)                      */
    int32_t is_input_backward;
    int32_t is_reserve_prox;  /* isReserveProximityGuard                                */
    int32_t shake;            /* spriteShakePosition                                    */
    int32_t has_won;
    int32_t input0;           /* input[0]: the input applied on the most recent frame   */
    uint32_t hist_left;       /* bit i = Left  held i frames ago (i = 0..31), from input[] */
    uint32_t hist_right;      /* bit i = Right held i frames ago                        */
    int32_t attack_run;       /* # consecutive most-recent frames with Attack held, saturated at 59 */
} fo_fighter_state;

/* Everything observable about one env after one call (reset or step). */
typedef struct {
    fo_fighter_state f[2];
    int32_t frame;            /* BattleCore.frameCount (-1 right after reset)           */
    int32_t recorded_input[2];/* p{1,2}MostRecentAction as GetEnvironmentState reports  */
    int32_t events;           /* bit0 P1's attack connected this frame, bit1 P2's; bits 2-3 P1-as-victim result
                                 (1 dmg,2 guard,3 break) , bits 4-5 P2-as-victim result; bit6 P1 prox-notified, bit7 P2 */
    int32_t battle_over;      /* game side: a fighter died on this frame                */
    int32_t was_reset;        /* this call performed the (auto-)reset sequence          */
    int32_t rng_draws;        /* total Random.Range calls consumed by this env so far   */
    uint32_t rng_state[4];
    int32_t bot_input[2];     /* the actor's held input for the NEXT frame (bot actors) */
    /* Python side (footsies.py:336-405, 518-570) */
    float obs[8];             /* guard p1,p2 | move idx p1,p2 | move_frame p1,p2 | position p1,p2 */
    float reward;
    int32_t terminated;
    int32_t info_frame;
    int32_t info_action[2];   /* bitmask L=1 R=2 A=4                                    */
    int32_t info_hitstun[2];
    double reward_f64;        /* the Python float before the fp32 cast                  */
} fo_trace;

typedef struct {
    int32_t p1_bot;           /* 1: P1 driven by BattleAI (reference by_example, --p1-bot)  */
    int32_t p2_bot;           /* 1: P2 driven by BattleAI (--p2-bot); 0: P2 from action tape */
    int32_t dense_reward;     /* footsies.py dense_reward                                */
    int32_t frame_delay;      /* footsies.py frame_delay                                 */
    int32_t autoreset;        /* 0 disabled (done envs freeze), 1 next-step (the call after a terminal
                                 step performs the reset and returns the frame -1 state)  */
    int32_t stale_intro_input;/* 1 (reference behaviour): the Intro frame replays the actors' last input */
} fo_config;

typedef struct fo_batch fo_batch;

fo_batch *fo_create(int32_t num_envs, const fo_config *cfg, int64_t first_env_index);
void fo_destroy(fo_batch *b);
/* Random.InitState(seed_base + global env index) for every env with mask (NULL = all). */
void fo_seed(fo_batch *b, int64_t seed_base, const uint8_t *mask);
/* Replace the RNG of one env by a tape of raw 32-bit draws (tests of the bot logic). */
void fo_set_rng_tape(fo_batch *b, int32_t env, const uint32_t *raw, int32_t n);
/* RESET command + read first state: envs with mask (NULL = all). out may be NULL. */
void fo_reset(fo_batch *b, const uint8_t *mask, fo_trace *out);
/* One FootsiesEnv.step per env.  actions: uint8 [num_envs] bitmask (L=1,R=2,A=4) for P1,
 * actions_p2 same for P2 (ignored / may be NULL when p2_bot).  repeat = frame-skip K: the same
 * action is applied for up to K frames, stopping at termination; reward is summed.
 * out (num_envs entries, may be NULL) receives the state after the last simulated frame. */
void fo_step(fo_batch *b, const uint8_t *actions_p1, const uint8_t *actions_p2, int32_t repeat,
             fo_trace *out, int32_t num_threads);
/* Overwrite the game state of one env (positions etc.), e.g. to set up KATs.  Only the fields of
 * fo_fighter_state plus frame are taken; history is rebuilt from hist_left/right/attack_run. */
void fo_set_state(fo_batch *b, int32_t env, const fo_fighter_state *p1, const fo_fighter_state *p2, int32_t frame);
void fo_get_trace(fo_batch *b, int32_t env, fo_trace *out);

/* Full battle state in the reference's save / load schema (BattleState.cs:9-24, FighterState.cs:26-56): the whole
 * 180-entry input arrays and the boxes as they stand.  fo_save_battle_state = BattleCore.SaveState
 * (BattleCore.cs:667-674, Fighter.cs:721-737); fo_load_battle_state = BattleCore.LoadState (BattleCore.cs:677-683,
 * Fighter.cs:739-811): fighters and frameCount only -- actors, bot queues, RNG and the Python-side state of the env
 * stay as they are. */
#define FO_INPUT_RECORD_FRAME 180
#define FO_MAX_BOXES 8
typedef struct { float x, y, width, height; } fo_rect;
typedef struct { fo_rect rect; int32_t proximity; int32_t attackID; } fo_hitbox;
typedef struct {
    float position[2];
    float velocity_x;
    int32_t isFaceRight;
    int32_t n_hitboxes; fo_hitbox hitboxes[FO_MAX_BOXES];
    int32_t n_hurtboxes; fo_rect hurtboxes[FO_MAX_BOXES];
    fo_rect pushbox;
    int32_t vitalHealth, guardHealth;
    int32_t currentActionID, currentActionFrame, currentActionHitCount, currentHitStunFrame;
    int32_t input[FO_INPUT_RECORD_FRAME], inputDown[FO_INPUT_RECORD_FRAME], inputUp[FO_INPUT_RECORD_FRAME];
    int32_t isInputBackward, isReserveProximityGuard;
    int32_t bufferActionID, reserveDamageActionID;
    int32_t spriteShakePosition, maxSpriteShakeFrame;
    int32_t hasWon;
} fo_full_fighter;
typedef struct { fo_full_fighter p[2]; float roundStartTime; int32_t frameCount; } fo_battle_state;
void fo_save_battle_state(fo_batch *b, int32_t env, fo_battle_state *out);
void fo_load_battle_state(fo_batch *b, int32_t env, const fo_battle_state *in);
/* Total fight frames simulated so far (the env-frames counter of the metric). */
int64_t fo_frames_simulated(fo_batch *b);
/* Episode statistics accumulated so far: see FO_STAT_* */
enum { FO_STAT_EPISODES = 0, FO_STAT_P1_WINS, FO_STAT_P2_WINS, FO_STAT_DOUBLE_KO, FO_STAT_FRAMES,
       FO_STAT_P1_SPECIALS, FO_STAT_P1_SPECIALS_NEUTRAL, FO_STAT_GUARD_BREAKS, FO_STAT_HITS, FO_STAT_BLOCKS,
       FO_STAT_COUNT };
void fo_stats(fo_batch *b, int64_t *out /* FO_STAT_COUNT */, double *return_sum);
/* Unity-style RNG helpers exposed for tests. */
void fo_rng_init(uint32_t s[4], int32_t seed);
uint32_t fo_rng_next(uint32_t s[4]);

#ifdef __cplusplus
}
#endif
#endif
