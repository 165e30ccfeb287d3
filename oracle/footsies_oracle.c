/*
 * footsies_oracle.c -- CPU ORACLE (test infrastructure, NOT product code; see footsies_oracle.h).
 *
 * Scalar restatement of the reference's per-frame battle update.  All file:line citations are
 * relative to /root/reference/.  Compile with -ffp-contract=off: every fp32 operation below is
 * meant to round separately, the way the C# float expressions are written.
 *
 * Parity: pinned, trace field by trace field, to the reference's own C# battle code transliterated mechanically into
 * C++ (oracle/_ref, tools/cs2cpp.py; tests/test_oracle_vs_ref.py), to hand-derived known answers and the reference's
 * moves.py table, and -- for the Python half -- to golden vectors produced by the reference's own FootsiesEnv
 * (tests/golden/).  Unpinned remains only third-party code outside /root/reference (UnityEngine.Random / Rect / Time, Mono
 * fp32 evaluation): no C# runtime or game binary exists offline.
 */
#include "footsies_oracle.h"
#include "frame_data.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

/* ---- InputDefine (InputData.cs:8-14) ---- */
enum { IN_LEFT = 1, IN_RIGHT = 2, IN_ATTACK = 4 };

/* ---- CommonActionID (Fighter.cs:42-61) ---- */
enum { A_STAND = 0, A_FORWARD = 1, A_BACKWARD = 2, A_DASH_FORWARD = 10, A_DASH_BACKWARD = 11,
       A_N_ATTACK = 100, A_B_ATTACK = 105, A_N_SPECIAL = 110, A_B_SPECIAL = 115, A_DAMAGE = 200,
       A_GUARD_M = 301, A_GUARD_STAND = 305, A_GUARD_CROUCH = 306, A_GUARD_BREAK = 310,
       A_GUARD_PROXIMITY = 350, A_DEAD = 500, A_WIN = 510 };
/* ActionType (ActionData.cs:61-67) */
enum { T_MOVEMENT = 0, T_ATTACK = 1, T_DAMAGE = 2, T_GUARD = 3 };
/* DamageResult (Fighter.cs:63-69) */
enum { DR_DAMAGE = 1, DR_GUARD = 2, DR_GUARD_BREAK = 3 };
/* RoundStateType (BattleCore.cs:13-20) */
enum { RS_STOP = 0, RS_INTRO, RS_FIGHT, RS_KO, RS_END };

#define INPUT_RECORD_FRAME 180          /* Fighter.cs:98 */
#define MAX_BOXES 8
#define MAX_RECORDING_INPUT_FRAME (60 * 60 * 5) /* BattleCore.cs:67 */
#define MAX_FIGHT_STATE_RECORD 10       /* BattleAI.cs:31 */
#define FIGHT_STATE_READ_INDEX 5        /* BattleAI.cs:32 */
#define QUEUE_CAP 512

typedef struct { float x, y, width, height; } Rect; /* UnityEngine.Rect: x,y = min corner */

typedef struct { Rect rect; int proximity; int attackID; } Hitbox;   /* Fighter.cs:28-32 */
typedef struct { Rect rect; } Box;                                   /* Hurtbox / Pushbox */

/* BoxBase accessors (Fighter.cs:12-15): x is treated as the CENTRE here */
static float box_xMin(const Rect *r) { return r->x - r->width / 2; }
static float box_xMax(const Rect *r) { return r->x + r->width / 2; }
static float box_yMin(const Rect *r) { return r->y; }
static float box_yMax(const Rect *r) { return r->y + r->height; }
/* BoxBase.Overlaps (Fighter.cs:17-25): inclusive */
static int box_overlaps(const Rect *self, const Rect *other) {
    int c1 = box_xMax(other) >= box_xMin(self);
    int c2 = box_xMin(other) <= box_xMax(self);
    int c3 = box_yMax(other) >= box_yMin(self);
    int c4 = box_yMin(other) <= box_yMax(self);
    return c1 && c2 && c3 && c4;
}
/* UnityEngine.Rect (engine, not in repo; documented behaviour): x is the LEFT edge, strict overlap */
static float urect_xMin(const Rect *r) { return r->x; }
static float urect_xMax(const Rect *r) { return r->width + r->x; }
static float urect_yMin(const Rect *r) { return r->y; }
static float urect_yMax(const Rect *r) { return r->height + r->y; }
static int urect_overlaps(const Rect *self, const Rect *other) {
    return urect_xMax(other) > urect_xMin(self) && urect_xMin(other) < urect_xMax(self)
        && urect_yMax(other) > urect_yMin(self) && urect_yMin(other) < urect_yMax(self);
}

typedef struct {
    float pos_x, pos_y;
    float velocity_x;
    int isFaceRight;
    int n_hitboxes; Hitbox hitboxes[MAX_BOXES];
    int n_hurtboxes; Box hurtboxes[MAX_BOXES];
    Box pushbox;
    int vitalHealth, guardHealth;
    int currentActionID, currentActionFrame, currentActionHitCount, currentHitStunFrame;
    int input[INPUT_RECORD_FRAME], inputDown[INPUT_RECORD_FRAME], inputUp[INPUT_RECORD_FRAME];
    int isInputBackward, isReserveProximityGuard;
    int bufferActionID, reserveDamageActionID;
    int spriteShakePosition, maxSpriteShakeFrame;
    int hasWon;
} Fighter;

typedef struct {
    float distanceX;
    int isOpponentDamage, isOpponentGuardBreak, isOpponentBlocking, isOpponentNormalAttack, isOpponentSpecialAttack;
    int isSet;   /* 0 = the slot still holds the C# array's initial null (BattleAI.cs:30) */
} FightState;

typedef struct { int buf[QUEUE_CAP]; int head, count; } Queue;
static void q_clear(Queue *q) { q->head = 0; q->count = 0; }
static void q_push(Queue *q, int v) { q->buf[(q->head + q->count) % QUEUE_CAP] = v; q->count++; }
static int q_pop(Queue *q) { int v = q->buf[q->head]; q->head = (q->head + 1) % QUEUE_CAP; q->count--; return v; }

typedef struct {
    int isPlayer1;
    Queue moveQueue, attackQueue;
    FightState fightStates[MAX_FIGHT_STATE_RECORD];
} BattleAI;

typedef struct {
    int p1Vital, p2Vital, p1Guard, p2Guard, p1Move, p1MoveFrame, p2Move, p2MoveFrame;
    float p1Position, p2Position;
    int globalFrame, p1MostRecentAction, p2MostRecentAction, p1Hitstun, p2Hitstun;
} EnvironmentState; /* EnvironmentState.cs:12-26 */

typedef struct {
    /* ---- game side ---- */
    Fighter fighter[2];
    BattleAI ai[2];
    int roundState;
    int frameCount;
    int actorInput[2];         /* TrainingRemoteActor.input / TrainingBattleAIActor.input: never cleared */
    unsigned currentRecordingInputIndex;
    int lastRecordedInput[2];  /* recordingP?Input[currentRecordingInputIndex-1].input */
    uint32_t rng[4];
    const uint32_t *tape; int tape_n, tape_pos;
    int rng_draws;
    int events;
    /* ---- python side (FootsiesEnv fields) ---- */
    EnvironmentState current_state;           /* self._current_state */
    EnvironmentState *delayed; int dq_len;    /* self.delayed_frame_queue (maxlen frame_delay+1) */
    double cumulative_episode_reward;          /* self._cummulative_episode_reward */
    int has_terminated;
    fo_trace last;
    /* ---- statistics ---- */
    int64_t stats[FO_STAT_COUNT];
    double return_sum, episode_return;
    int episode_frames;
    int64_t frames_simulated;
} Env;

struct fo_batch {
    int n;
    fo_config cfg;
    int64_t first_env_index;
    Env *envs;
};

/* ------------------------------------------------------------------------------------------
 * Frame data lookups (ActionData.cs:87-168): linear scans, inclusive ranges
 * ---------------------------------------------------------------------------------------- */
static const fd_action *get_action(int id) {
    for (int i = 0; i < FD_NUM_ACTIONS; i++) if (FD_ACTIONS[i].actionID == id) return &FD_ACTIONS[i];
    return &FD_ACTIONS[0];
}
static const fd_attack *get_attack(int id) {
    for (int i = 0; i < FD_NUM_ATTACKS; i++) if (FD_ATTACKS[i].attackID == id) return &FD_ATTACKS[i];
    return 0;
}
static int move_index(int id) { /* moves.py:41-42 FOOTSIES_MOVE_ID_TO_INDEX */
    for (int i = 0; i < FD_NUM_ACTIONS; i++) if (FD_ACTIONS[i].actionID == id) return i;
    return 0;
}
/* ActionData.GetMovementData (ActionData.cs:146-155): FIRST match */
static const fd_movement *get_movement(const fd_action *a, int frame) {
    for (int i = 0; i < a->n_movements; i++)
        if (frame >= a->movements[i].start && frame <= a->movements[i].end) return &a->movements[i];
    return 0;
}
/* ActionData.GetPushboxData (ActionData.cs:135-144): FIRST match */
static const fd_box *get_pushbox(const fd_action *a, int frame) {
    for (int i = 0; i < a->n_pushboxes; i++)
        if (frame >= a->pushboxes[i].start && frame <= a->pushboxes[i].end) return &a->pushboxes[i];
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * UnityEngine.Random (engine; restated from public descriptions, UNPINNED): xorshift128
 * ---------------------------------------------------------------------------------------- */
void fo_rng_init(uint32_t s[4], int32_t seed) { /* Random.InitState(int) */
    s[0] = (uint32_t)seed;
    s[1] = s[0] * 1812433253u + 1u;
    s[2] = s[1] * 1812433253u + 1u;
    s[3] = s[2] * 1812433253u + 1u;
}
uint32_t fo_rng_next(uint32_t s[4]) {
    uint32_t t = s[0] ^ (s[0] << 11);
    s[0] = s[1]; s[1] = s[2]; s[2] = s[3];
    s[3] = s[3] ^ (s[3] >> 19) ^ t ^ (t >> 8);
    return s[3];
}
/* Random.Range(int minInclusive, int maxExclusive) */
static int random_range(Env *e, int min, int max) {
    uint32_t r;
    if (e->tape) { r = e->tape_pos < e->tape_n ? e->tape[e->tape_pos] : 0; e->tape_pos++; }
    else r = fo_rng_next(e->rng);
    e->rng_draws++;
    return min + (int)(r % (uint32_t)(max - min));
}

/* ------------------------------------------------------------------------------------------
 * Fighter (Fighter.cs)
 * ---------------------------------------------------------------------------------------- */
static int f_isActionEnd(const Fighter *f) { /* Fighter.cs:90 */
    return f->currentActionFrame >= get_action(f->currentActionID)->frameCount;
}
static int f_isDead(const Fighter *f) { return f->vitalHealth <= 0; } /* Fighter.cs:83 */

static int IsAttackInput(int input) { return (input & IN_ATTACK) > 0; } /* Fighter.cs:637 */
static int IsForwardInput(const Fighter *f, int input) {                 /* Fighter.cs:642-653 */
    return f->isFaceRight ? (input & IN_RIGHT) > 0 : (input & IN_LEFT) > 0;
}
static int IsBackwardInput(const Fighter *f, int input) {                /* Fighter.cs:655-666 */
    return f->isFaceRight ? (input & IN_LEFT) > 0 : (input & IN_RIGHT) > 0;
}

/* Fighter.SetCurrentAction (Fighter.cs:546-563); audio omitted (no state effect) */
static void f_SetCurrentAction(Fighter *f, int actionID, int startFrame) {
    f->currentActionID = actionID;
    f->currentActionFrame = startFrame;
    f->currentActionHitCount = 0;
    f->bufferActionID = -1;
    f->reserveDamageActionID = -1;
    f->spriteShakePosition = 0;
}

static void f_ClearInput(Fighter *f) { /* Fighter.cs:521-529 */
    for (int i = 0; i < INPUT_RECORD_FRAME; i++) { f->input[i] = 0; f->inputDown[i] = 0; f->inputUp[i] = 0; }
}

/* Fighter.SetupBattleStart (Fighter.cs:120-135) */
static void f_SetupBattleStart(Fighter *f, float start_x, int isPlayerOne) {
    f->pos_x = start_x; f->pos_y = 0.0f;
    f->isFaceRight = isPlayerOne;
    f->vitalHealth = 1;
    f->guardHealth = FD_START_GUARD_HEALTH;
    f->hasWon = 0;
    f->velocity_x = 0;
    f_ClearInput(f);
    f_SetCurrentAction(f, A_STAND, 0);
}

/* Fighter.IncrementActionFrame (Fighter.cs:140-166) */
static void f_IncrementActionFrame(Fighter *f) {
    if (abs(f->spriteShakePosition) > 0) {
        f->spriteShakePosition *= -1;
        f->spriteShakePosition += (f->spriteShakePosition > 0 ? -1 : 1);
    }
    if (f->currentHitStunFrame > 0) {
        f->currentHitStunFrame--;
        return;
    }
    f->currentActionFrame++;
    if (f_isActionEnd(f)) {
        const fd_action *a = get_action(f->currentActionID);
        if (a->isLoop) f->currentActionFrame = a->loopFromFrame;
    }
}

/* Fighter.UpdateInput (Fighter.cs:172-188) */
static void f_UpdateInput(Fighter *f, int inputData) {
    for (int i = INPUT_RECORD_FRAME - 1; i >= 1; i--) {
        f->input[i] = f->input[i - 1];
        f->inputDown[i] = f->inputDown[i - 1];
        f->inputUp[i] = f->inputUp[i - 1];
    }
    f->input[0] = inputData;
    f->inputDown[0] = (f->input[0] ^ f->input[1]) & f->input[0];
    f->inputUp[0] = (f->input[0] ^ f->input[1]) & ~f->input[0];
}

/* Fighter.RequestAction (Fighter.cs:472-510) */
static int f_RequestAction(Fighter *f, int actionID) {
    if (f_isActionEnd(f)) {
        f_SetCurrentAction(f, actionID, 0);
        return 1;
    }
    if (f->currentActionID == actionID) return 0;
    const fd_action *a = get_action(f->currentActionID);
    if (a->alwaysCancelable) {
        f_SetCurrentAction(f, actionID, 0);
        return 1;
    } else {
        /* ActionData.GetCancelData (ActionData.cs:157-168): ALL matches, in order */
        for (int i = 0; i < a->n_cancels; i++) {
            const fd_cancel *c = &a->cancels[i];
            if (!(f->currentActionFrame >= c->start && f->currentActionFrame <= c->end)) continue;
            int contains = 0;
            for (int k = 0; k < c->n_ids; k++) if (c->ids[k] == actionID) contains = 1;
            if (contains) {
                if (c->execute) { f->bufferActionID = actionID; return 1; }
                else if (c->buffer) { f->bufferActionID = actionID; }
            }
        }
    }
    return 0;
}

/* Fighter.CheckSpecialAttackInput (Fighter.cs:569-583) */
static int f_CheckSpecialAttackInput(const Fighter *f) {
    if (!IsAttackInput(f->inputUp[0])) return 0;
    for (int i = 1; i < FD_SPECIAL_ATTACK_HOLD_FRAME; i++)
        if (!IsAttackInput(f->input[i])) return 0;
    return 1;
}
/* Fighter.CheckForwardDashInput (Fighter.cs:585-609) */
static int f_CheckForwardDashInput(const Fighter *f) {
    if (!IsForwardInput(f, f->inputDown[0])) return 0;
    for (int i = 1; i < FD_DASH_ALLOW_FRAME; i++) {
        if (IsBackwardInput(f, f->input[i])) return 0;
        if (IsForwardInput(f, f->input[i])) {
            for (int j = i + 1; j < i + FD_DASH_ALLOW_FRAME; j++)
                if (!IsForwardInput(f, f->input[j]) && !IsBackwardInput(f, f->input[j])) return 1;
            return 0;
        }
    }
    return 0;
}
/* Fighter.CheckBackwardDashInput (Fighter.cs:611-635) */
static int f_CheckBackwardDashInput(const Fighter *f) {
    if (!IsBackwardInput(f, f->inputDown[0])) return 0;
    for (int i = 1; i < FD_DASH_ALLOW_FRAME; i++) {
        if (IsForwardInput(f, f->input[i])) return 0;
        if (IsBackwardInput(f, f->input[i])) {
            for (int j = i + 1; j < i + FD_DASH_ALLOW_FRAME; j++)
                if (!IsForwardInput(f, f->input[j]) && !IsBackwardInput(f, f->input[j])) return 1;
            return 0;
        }
    }
    return 0;
}

static int f_canCancelAttack(const Fighter *f) { /* Fighter.cs:531-539 */
    if (FD_CAN_CANCEL_ON_WHIFF) return 1;
    else if (f->currentActionHitCount > 0) return 1;
    return 0;
}

/* Fighter.UpdateActionRequest (Fighter.cs:201-286) */
static void f_UpdateActionRequest(Fighter *f) {
    if (f->hasWon) { f_RequestAction(f, A_WIN); return; }

    if (f->reserveDamageActionID != -1 && f->currentHitStunFrame <= 0) {
        f_SetCurrentAction(f, f->reserveDamageActionID, 0);
        f->reserveDamageActionID = -1;
        return;
    }
    if (f->bufferActionID != -1 && f_canCancelAttack(f) && f->currentHitStunFrame <= 0) {
        f_SetCurrentAction(f, f->bufferActionID, 0);
        f->bufferActionID = -1;
        return;
    }

    int isForward = IsForwardInput(f, f->input[0]);
    int isBackward = IsBackwardInput(f, f->input[0]);
    int isAttack = IsAttackInput(f->inputDown[0]);
    if (f_CheckSpecialAttackInput(f)) {
        if (isBackward || isForward) f_RequestAction(f, A_B_SPECIAL);
        else f_RequestAction(f, A_N_SPECIAL);
    } else if (isAttack) {
        if ((f->currentActionID == A_N_ATTACK || f->currentActionID == A_B_ATTACK) && !f_isActionEnd(f))
            f_RequestAction(f, A_N_SPECIAL);
        else {
            if (isBackward || isForward) f_RequestAction(f, A_B_ATTACK);
            else f_RequestAction(f, A_N_ATTACK);
        }
    }

    if (f_CheckForwardDashInput(f)) f_RequestAction(f, A_DASH_FORWARD);
    else if (f_CheckBackwardDashInput(f)) f_RequestAction(f, A_DASH_BACKWARD);

    f->isInputBackward = isBackward;

    if (isForward && isBackward) f_RequestAction(f, A_STAND);
    else if (isForward) f_RequestAction(f, A_FORWARD);
    else if (isBackward) {
        if (f->isReserveProximityGuard) f_RequestAction(f, A_GUARD_PROXIMITY);
        else f_RequestAction(f, A_BACKWARD);
    } else f_RequestAction(f, A_STAND);

    f->isReserveProximityGuard = 0;
}

/* Fighter.UpdateMovement (Fighter.cs:291-319); Time.deltaTime inside FixedUpdate == fixedDeltaTime */
static void f_UpdateMovement(Fighter *f) {
    if (f->currentHitStunFrame > 0) return;
    int sign = f->isFaceRight ? 1 : -1;
    const float dt = FD_FIXED_DELTA_TIME;
    if (f->currentActionID == A_FORWARD) {
        f->pos_x += FD_FORWARD_MOVE_SPEED * sign * dt;
        return;
    } else if (f->currentActionID == A_BACKWARD) {
        f->pos_x -= FD_BACKWARD_MOVE_SPEED * sign * dt;
        return;
    }
    const fd_movement *m = get_movement(get_action(f->currentActionID), f->currentActionFrame);
    if (m) {
        f->velocity_x = m->velocity_x;
        if (f->velocity_x != 0) f->pos_x += f->velocity_x * sign * dt;
    }
}

/* Fighter.TransformToFightRect (Fighter.cs:706-719) */
static Rect TransformToFightRect(fd_rect d, float bx, float by, int isFaceRight) {
    int sign = isFaceRight ? 1 : -1;
    Rect r;
    r.x = bx + (d.x * sign);
    r.y = by + d.y;
    r.width = d.width;
    r.height = d.height;
    return r;
}

/* Fighter.UpdateBoxes -> ApplyCurrentActionData (Fighter.cs:321-324, 671-697) */
static void f_UpdateBoxes(Fighter *f) {
    const fd_action *a = get_action(f->currentActionID);
    int frame = f->currentActionFrame;
    f->n_hitboxes = 0;
    f->n_hurtboxes = 0;
    for (int i = 0; i < a->n_hitboxes; i++) { /* GetHitboxData: ALL matches (ActionData.cs:109-120) */
        const fd_hitbox *h = &a->hitboxes[i];
        if (frame >= h->start && frame <= h->end) {
            Hitbox *b = &f->hitboxes[f->n_hitboxes++];
            b->rect = TransformToFightRect(h->rect, f->pos_x, f->pos_y, f->isFaceRight);
            b->proximity = h->proximity;
            b->attackID = h->attackID;
        }
    }
    for (int i = 0; i < a->n_hurtboxes; i++) { /* GetHurtboxData: ALL matches (ActionData.cs:122-133) */
        const fd_box *h = &a->hurtboxes[i];
        if (frame >= h->start && frame <= h->end) {
            fd_rect r = h->useBaseRect ? FD_BASE_HURTBOX : h->rect;
            f->hurtboxes[f->n_hurtboxes++].rect = TransformToFightRect(r, f->pos_x, f->pos_y, f->isFaceRight);
        }
    }
    const fd_box *p = get_pushbox(a, frame);
    /* the reference would NRE on a missing pushbox; every reachable (action, frame) has one */
    fd_rect pr = (p == 0 || p->useBaseRect) ? FD_BASE_PUSHBOX : p->rect;
    f->pushbox.rect = TransformToFightRect(pr, f->pos_x, f->pos_y, f->isFaceRight);
}

/* Fighter.ApplyPositionChange (Fighter.cs:331-350) */
static void f_ApplyPositionChange(Fighter *f, float x, float y) {
    f->pos_x += x; f->pos_y += y;
    for (int i = 0; i < f->n_hitboxes; i++) { f->hitboxes[i].rect.x += x; f->hitboxes[i].rect.y += y; }
    for (int i = 0; i < f->n_hurtboxes; i++) { f->hurtboxes[i].rect.x += x; f->hurtboxes[i].rect.y += y; }
    f->pushbox.rect.x += x; f->pushbox.rect.y += y;
}

/* Fighter.NotifyDamaged (Fighter.cs:357-398) */
static int f_NotifyDamaged(Fighter *f, const fd_attack *atk) {
    int isGuardBreak = 0;
    if (atk->guardHealthDamage > 0) {
        f->guardHealth -= atk->guardHealthDamage;
        if (f->guardHealth < 0) { isGuardBreak = 1; f->guardHealth = 0; }
    }
    if (f->currentActionID == A_BACKWARD || get_action(f->currentActionID)->type == T_GUARD) {
        if (isGuardBreak) {
            f_SetCurrentAction(f, atk->guardActionID, 0);
            f->reserveDamageActionID = A_GUARD_BREAK;
            return DR_GUARD_BREAK;
        } else {
            f_SetCurrentAction(f, atk->guardActionID, 0);
            return DR_GUARD;
        }
    } else {
        if (atk->vitalHealthDamage > 0) {
            f->vitalHealth -= atk->vitalHealthDamage;
            if (f->vitalHealth <= 0) f->vitalHealth = 0;
        }
        f_SetCurrentAction(f, atk->damageActionID, 0);
        return DR_DAMAGE;
    }
}

static int f_CanAttackHit(const Fighter *f, int attackID) { /* Fighter.cs:408-420 */
    const fd_attack *a = get_attack(attackID);
    if (!a) return 1;
    if (f->currentActionHitCount >= a->numberOfHit) return 0;
    return 1;
}
static int f_GetHitStunFrame(int damageResult, int attackID) { /* Fighter.cs:446-454 */
    const fd_attack *a = get_attack(attackID);
    if (damageResult == DR_GUARD) return a->guardStunFrame;
    else if (damageResult == DR_GUARD_BREAK) return a->guardBreakStunFrame;
    return a->hitStunFrame;
}
static void f_SetSpriteShakeFrame(Fighter *f, int spriteShakeFrame) { /* Fighter.cs:438-444 */
    if (spriteShakeFrame > f->maxSpriteShakeFrame) spriteShakeFrame = f->maxSpriteShakeFrame;
    f->spriteShakePosition = spriteShakeFrame * (f->isFaceRight ? -1 : 1);
}

/* ------------------------------------------------------------------------------------------
 * BattleAI (BattleAI.cs)
 * ---------------------------------------------------------------------------------------- */
static int ai_GetForwardInput(const BattleAI *ai) { return ai->isPlayer1 ? IN_RIGHT : IN_LEFT; }  /* :380 */
static int ai_GetBackwardInput(const BattleAI *ai) { return ai->isPlayer1 ? IN_LEFT : IN_RIGHT; } /* :385 */
static void ai_AddForwardInputQueue(BattleAI *ai, int n) { for (int i = 0; i < n; i++) q_push(&ai->moveQueue, ai_GetForwardInput(ai)); }
static void ai_AddBackwardInputQueue(BattleAI *ai, int n) { for (int i = 0; i < n; i++) q_push(&ai->moveQueue, ai_GetBackwardInput(ai)); }
static void ai_AddForwardDashInputQueue(BattleAI *ai) { /* :330-335 */
    q_push(&ai->moveQueue, ai_GetForwardInput(ai)); q_push(&ai->moveQueue, 0); q_push(&ai->moveQueue, ai_GetForwardInput(ai));
}
static void ai_AddBackwardDashInputQueue(BattleAI *ai) { /* :337-342 -- enqueues FORWARD taps (reference bug, kept) */
    q_push(&ai->moveQueue, ai_GetForwardInput(ai)); q_push(&ai->moveQueue, 0); q_push(&ai->moveQueue, ai_GetForwardInput(ai));
}
static void ai_AddNeutralMovement(BattleAI *ai) { for (int i = 0; i < 30; i++) q_push(&ai->moveQueue, 0); } /* :192 */
static void ai_AddFarApproach1(BattleAI *ai) { ai_AddForwardInputQueue(ai, 40); ai_AddBackwardInputQueue(ai, 10); ai_AddForwardInputQueue(ai, 30); ai_AddBackwardInputQueue(ai, 10); }
static void ai_AddFarApproach2(BattleAI *ai) { ai_AddForwardDashInputQueue(ai); ai_AddBackwardInputQueue(ai, 25); ai_AddForwardDashInputQueue(ai); ai_AddBackwardInputQueue(ai, 25); }
static void ai_AddMidApproach1(BattleAI *ai) { ai_AddForwardInputQueue(ai, 30); ai_AddBackwardInputQueue(ai, 10); ai_AddForwardInputQueue(ai, 20); ai_AddBackwardInputQueue(ai, 10); }
static void ai_AddMidApproach2(BattleAI *ai) { ai_AddForwardDashInputQueue(ai); ai_AddBackwardInputQueue(ai, 30); }
static void ai_AddFallBack1(BattleAI *ai) { ai_AddBackwardInputQueue(ai, 60); }
static void ai_AddFallBack2(BattleAI *ai) { ai_AddBackwardDashInputQueue(ai); ai_AddBackwardInputQueue(ai, 60); }
static void ai_AddNoAttack(BattleAI *ai) { for (int i = 0; i < 30; i++) q_push(&ai->attackQueue, 0); }
static void ai_AddOneHitImmediateAttack(BattleAI *ai) { q_push(&ai->attackQueue, IN_ATTACK); for (int i = 0; i < 18; i++) q_push(&ai->attackQueue, 0); }
static void ai_AddTwoHitImmediateAttack(BattleAI *ai) {
    q_push(&ai->attackQueue, IN_ATTACK); for (int i = 0; i < 3; i++) q_push(&ai->attackQueue, 0);
    q_push(&ai->attackQueue, IN_ATTACK); for (int i = 0; i < 18; i++) q_push(&ai->attackQueue, 0);
}
static void ai_AddImmediateSpecialAttack(BattleAI *ai) { for (int i = 0; i < 60; i++) q_push(&ai->attackQueue, IN_ATTACK); q_push(&ai->attackQueue, 0); }
static void ai_AddDelaySpecialAttack(BattleAI *ai) { for (int i = 0; i < 120; i++) q_push(&ai->attackQueue, IN_ATTACK); q_push(&ai->attackQueue, 0); }

/* BattleAI.SelectMovement (BattleAI.cs:68-126) */
static void ai_SelectMovement(Env *e, BattleAI *ai, const FightState *s) {
    if (s->distanceX > 4.0f) {
        int r = random_range(e, 0, 2);
        if (r == 0) ai_AddFarApproach1(ai); else ai_AddFarApproach2(ai);
    } else if (s->distanceX > 3.0f) {
        int r = random_range(e, 0, 7);
        if (r <= 1) ai_AddMidApproach1(ai);
        else if (r <= 3) ai_AddMidApproach2(ai);
        else if (r == 4) ai_AddFarApproach1(ai);
        else if (r == 5) ai_AddFarApproach2(ai);
        else ai_AddNeutralMovement(ai);
    } else if (s->distanceX > 2.5f) {
        int r = random_range(e, 0, 5);
        if (r == 0) ai_AddMidApproach1(ai);
        else if (r == 1) ai_AddMidApproach2(ai);
        else if (r == 2) ai_AddFallBack1(ai);
        else if (r == 3) ai_AddFallBack2(ai);
        else ai_AddNeutralMovement(ai);
    } else if (s->distanceX > 2.0f) {
        int r = random_range(e, 0, 4);
        if (r == 0) ai_AddFallBack1(ai);
        else if (r == 1) ai_AddFallBack2(ai);
        else ai_AddNeutralMovement(ai);
    } else {
        int r = random_range(e, 0, 3);
        if (r == 0) ai_AddFallBack1(ai);
        else if (r == 1) ai_AddFallBack2(ai);
        else ai_AddNeutralMovement(ai);
    }
}
/* BattleAI.SelectAttack (BattleAI.cs:128-190) */
static void ai_SelectAttack(Env *e, BattleAI *ai, const FightState *s) {
    if (s->isOpponentDamage || s->isOpponentGuardBreak || s->isOpponentSpecialAttack) {
        ai_AddTwoHitImmediateAttack(ai);
    } else if (s->distanceX > 4.0f) {
        int r = random_range(e, 0, 4);
        if (r <= 3) ai_AddNoAttack(ai); else ai_AddDelaySpecialAttack(ai);
    } else if (s->distanceX > 3.0f) {
        if (s->isOpponentNormalAttack) { ai_AddTwoHitImmediateAttack(ai); return; }
        int r = random_range(e, 0, 5);
        if (r <= 1) ai_AddNoAttack(ai);
        else if (r <= 3) ai_AddOneHitImmediateAttack(ai);
        else ai_AddDelaySpecialAttack(ai);
    } else if (s->distanceX > 2.5f) {
        int r = random_range(e, 0, 3);
        if (r == 0) ai_AddNoAttack(ai);
        else if (r == 1) ai_AddOneHitImmediateAttack(ai);
        else ai_AddTwoHitImmediateAttack(ai);
    } else if (s->distanceX > 2.0f) {
        int r = random_range(e, 0, 6);
        if (r <= 1) ai_AddOneHitImmediateAttack(ai);
        else if (r <= 3) ai_AddTwoHitImmediateAttack(ai);
        else if (r == 4) ai_AddImmediateSpecialAttack(ai);
        else ai_AddDelaySpecialAttack(ai);
    } else {
        int r = random_range(e, 0, 3);
        if (r == 0) ai_AddOneHitImmediateAttack(ai);
        else ai_AddTwoHitImmediateAttack(ai);
    }
}
/* BattleAI.UpdateFightState (BattleAI.cs:344-363) incl. the ascending shift loop (every slot >= 1
 * becomes the PREVIOUS call's slot 0) */
static void ai_UpdateFightState(Env *e, BattleAI *ai) {
    const Fighter *opp = ai->isPlayer1 ? &e->fighter[1] : &e->fighter[0];
    FightState cur;
    cur.distanceX = fabsf(e->fighter[1].pos_x - e->fighter[0].pos_x);
    cur.isOpponentDamage = opp->currentActionID == A_DAMAGE;
    cur.isOpponentGuardBreak = opp->currentActionID == A_GUARD_BREAK;
    cur.isOpponentBlocking = (opp->currentActionID == A_GUARD_CROUCH || opp->currentActionID == A_GUARD_STAND
                              || opp->currentActionID == A_GUARD_M);
    cur.isOpponentNormalAttack = (opp->currentActionID == A_N_ATTACK || opp->currentActionID == A_B_ATTACK);
    cur.isOpponentSpecialAttack = (opp->currentActionID == A_N_SPECIAL || opp->currentActionID == A_B_SPECIAL);
    cur.isSet = 1;
    for (int i = 1; i < MAX_FIGHT_STATE_RECORD; i++) ai->fightStates[i] = ai->fightStates[i - 1];
    ai->fightStates[0] = cur;
}
/* BattleAI.getNextAIInput (BattleAI.cs:41-66) */
static int ai_getNextAIInput(Env *e, BattleAI *ai) {
    int input = 0;
    ai_UpdateFightState(e, ai);
    const FightState *s = &ai->fightStates[FIGHT_STATE_READ_INDEX];
    if (!s->isSet) return 0;                                   /* `if (fightState != null)` :47 -- only a BattleAI that
                                                                  was never Reset() sees this, on its very first call */
    if (ai->moveQueue.count > 0) input |= q_pop(&ai->moveQueue);
    else ai_SelectMovement(e, ai, s);
    if (ai->attackQueue.count > 0) input |= q_pop(&ai->attackQueue);
    else ai_SelectAttack(e, ai, s);
    return input;
}
/* BattleAI.Reset (BattleAI.cs:393-403) */
static void ai_Reset(Env *e, BattleAI *ai) {
    q_clear(&ai->moveQueue);
    q_clear(&ai->attackQueue);
    ai_UpdateFightState(e, ai);
    for (int i = 1; i < MAX_FIGHT_STATE_RECORD; i++) ai->fightStates[i] = ai->fightStates[0];
}

/* ------------------------------------------------------------------------------------------
 * BattleCore (BattleCore.cs)
 * ---------------------------------------------------------------------------------------- */
/* BattleCore.UpdatePushCharacterVsCharacter (BattleCore.cs:483-501): UnityEngine.Rect semantics */
static void UpdatePushCharacterVsCharacter(Env *e) {
    Fighter *f1 = &e->fighter[0], *f2 = &e->fighter[1];
    Rect rect1 = f1->pushbox.rect, rect2 = f2->pushbox.rect; /* struct copies, as in C# */
    if (urect_overlaps(&rect1, &rect2)) {
        if (f1->pos_x < f2->pos_x) {
            f_ApplyPositionChange(f1, (urect_xMax(&rect1) - urect_xMin(&rect2)) * -1 / 2, f1->pos_y);
            f_ApplyPositionChange(f2, (urect_xMax(&rect1) - urect_xMin(&rect2)) * 1 / 2, f2->pos_y);
        } else if (f1->pos_x > f2->pos_x) {
            f_ApplyPositionChange(f1, (urect_xMax(&rect2) - urect_xMin(&rect1)) * 1 / 2, f1->pos_y);
            f_ApplyPositionChange(f2, (urect_xMax(&rect2) - urect_xMin(&rect1)) * -1 / 2, f1->pos_y);
        }
    }
}
/* BattleCore.UpdatePushCharacterVsBackground (BattleCore.cs:503-519): BoxBase semantics */
static void UpdatePushCharacterVsBackground(Env *e) {
    float stageMinX = FD_BATTLE_AREA_WIDTH * -1 / 2;
    float stageMaxX = FD_BATTLE_AREA_WIDTH / 2;
    for (int i = 0; i < 2; i++) {
        Fighter *f = &e->fighter[i];
        if (box_xMin(&f->pushbox.rect) < stageMinX)
            f_ApplyPositionChange(f, stageMinX - box_xMin(&f->pushbox.rect), f->pos_y);
        else if (box_xMax(&f->pushbox.rect) > stageMaxX)
            f_ApplyPositionChange(f, stageMaxX - box_xMax(&f->pushbox.rect), f->pos_y);
    }
}
/* BattleCore.UpdateHitboxHurtboxCollision (BattleCore.cs:521-591) */
static void UpdateHitboxHurtboxCollision(Env *e) {
    for (int ai = 0; ai < 2; ai++) {
        Fighter *attacker = &e->fighter[ai];
        int isHit = 0, isProximity = 0, hitAttackID = 0;
        for (int di = 0; di < 2; di++) {
            if (di == ai) continue;
            Fighter *damaged = &e->fighter[di];
            for (int h = 0; h < attacker->n_hitboxes; h++) {
                const Hitbox *hitbox = &attacker->hitboxes[h];
                if (!f_CanAttackHit(attacker, hitbox->attackID)) continue;
                for (int u = 0; u < damaged->n_hurtboxes; u++) {
                    if (box_overlaps(&hitbox->rect, &damaged->hurtboxes[u].rect)) {
                        if (hitbox->proximity) isProximity = 1;
                        else { isHit = 1; hitAttackID = hitbox->attackID; break; }
                    }
                }
                if (isHit) break;
            }
            if (isHit) {
                attacker->currentActionHitCount++;                               /* NotifyAttackHit :352 */
                int vital_before = damaged->vitalHealth;
                int res = f_NotifyDamaged(damaged, get_attack(hitAttackID));
                int stun = f_GetHitStunFrame(res, hitAttackID);
                attacker->currentHitStunFrame = stun;                           /* SetHitStun :433 */
                damaged->currentHitStunFrame = stun;
                f_SetSpriteShakeFrame(damaged, stun / 3);
                e->events |= 1 << ai;
                e->events |= res << (2 + 2 * di);
                if (res == DR_GUARD_BREAK) e->stats[FO_STAT_GUARD_BREAKS]++;
                else if (res == DR_GUARD) e->stats[FO_STAT_BLOCKS]++;
                else e->stats[FO_STAT_HITS]++;
                (void)vital_before;
            } else if (isProximity) {
                if (damaged->isInputBackward) damaged->isReserveProximityGuard = 1; /* NotifyInProximityGuardRange :400 */
                e->events |= 1 << (6 + di);
            }
        }
    }
}

/* BattleCore.RecordInput (BattleCore.cs:593-607) */
static void RecordInput(Env *e, int p1, int p2) {
    if (e->currentRecordingInputIndex >= MAX_RECORDING_INPUT_FRAME) return;
    e->lastRecordedInput[0] = p1; e->lastRecordedInput[1] = p2;
    e->currentRecordingInputIndex++;
}
/* BattleCore.GetEnvironmentState (BattleCore.cs:449-468) */
static EnvironmentState GetEnvironmentState(const Env *e) {
    EnvironmentState s;
    const Fighter *f1 = &e->fighter[0], *f2 = &e->fighter[1];
    s.p1Vital = f1->vitalHealth; s.p2Vital = f2->vitalHealth;
    s.p1Guard = f1->guardHealth; s.p2Guard = f2->guardHealth;
    s.p1Move = f1->currentActionID; s.p1MoveFrame = f1->currentActionFrame;
    s.p2Move = f2->currentActionID; s.p2MoveFrame = f2->currentActionFrame;
    s.p1Position = f1->pos_x; s.p2Position = f2->pos_x;
    s.globalFrame = e->frameCount;
    s.p1MostRecentAction = e->currentRecordingInputIndex > 0 ? e->lastRecordedInput[0] : 0;
    s.p2MostRecentAction = e->currentRecordingInputIndex > 0 ? e->lastRecordedInput[1] : 0;
    s.p1Hitstun = f1->currentHitStunFrame; s.p2Hitstun = f2->currentHitStunFrame;
    return s;
}

/* TrainingManager.Step (TrainingManager.cs:59-77): bot actors are queried AFTER the frame, P1 first,
 * and not at all on the terminal frame; remote actors keep their last input until a new one arrives. */
static void TrainingManager_Step(Env *e, const fo_config *cfg, int battleOver) {
    if (!battleOver) {
        if (cfg->p1_bot) e->actorInput[0] = ai_getNextAIInput(e, &e->ai[0]); /* TrainingBattleAIActor.cs:38-41 */
        if (cfg->p2_bot) e->actorInput[1] = ai_getNextAIInput(e, &e->ai[1]);
    }
}

/* BattleCore.UpdateIntroState (BattleCore.cs:329-345) */
static void UpdateIntroState(Env *e) {
    int p1 = e->actorInput[0], p2 = e->actorInput[1];
    RecordInput(e, p1, p2);
    f_UpdateInput(&e->fighter[0], p1);
    f_UpdateInput(&e->fighter[1], p2);
    for (int i = 0; i < 2; i++) f_IncrementActionFrame(&e->fighter[i]);
    for (int i = 0; i < 2; i++) f_RequestAction(&e->fighter[i], A_STAND); /* UpdateIntroAction :193 */
    for (int i = 0; i < 2; i++) f_UpdateMovement(&e->fighter[i]);
    for (int i = 0; i < 2; i++) f_UpdateBoxes(&e->fighter[i]);
    UpdatePushCharacterVsCharacter(e);
    UpdatePushCharacterVsBackground(e);
}
/* BattleCore.UpdateFightState (BattleCore.cs:347-364) */
static void UpdateFightState(Env *e) {
    int p1 = e->actorInput[0], p2 = e->actorInput[1];
    RecordInput(e, p1, p2);
    f_UpdateInput(&e->fighter[0], p1);
    f_UpdateInput(&e->fighter[1], p2);
    for (int i = 0; i < 2; i++) f_IncrementActionFrame(&e->fighter[i]);
    for (int i = 0; i < 2; i++) f_UpdateActionRequest(&e->fighter[i]);
    for (int i = 0; i < 2; i++) f_UpdateMovement(&e->fighter[i]);
    for (int i = 0; i < 2; i++) f_UpdateBoxes(&e->fighter[i]);
    UpdatePushCharacterVsCharacter(e);
    UpdatePushCharacterVsBackground(e);
    UpdateHitboxHurtboxCollision(e);
}
/* BattleCore.UpdateEndState (BattleCore.cs:371-381) -- never observable in training, kept for fidelity */
static void UpdateEndState(Env *e) {
    for (int i = 0; i < 2; i++) f_IncrementActionFrame(&e->fighter[i]);
    for (int i = 0; i < 2; i++) f_UpdateActionRequest(&e->fighter[i]);
    for (int i = 0; i < 2; i++) f_UpdateMovement(&e->fighter[i]);
    for (int i = 0; i < 2; i++) f_UpdateBoxes(&e->fighter[i]);
    UpdatePushCharacterVsCharacter(e);
    UpdatePushCharacterVsBackground(e);
}

/* The FixedUpdates between a terminal Fight frame and the next Intro (BattleCore.cs:221-243, 292-326):
 * KO (ClearInput) -> End (RequestWinAction) -> one UpdateEndState -> Stop. */
static void run_ko_end(Env *e) {
    /* ChangeRoundState(KO) already executed on the terminal frame (:216, :292-305) */
    e->roundState = RS_END;                                   /* KO tick: timer 0 -> End (:221-230) */
    int dead = f_isDead(&e->fighter[0]) + f_isDead(&e->fighter[1]);
    if (dead == 1) {                                          /* :310-323 */
        if (f_isDead(&e->fighter[0])) e->fighter[1].hasWon = 1; else e->fighter[0].hasWon = 1;
    }
    UpdateEndState(e);                                        /* End tick (:232-243) */
    e->roundState = RS_STOP;
}

/* Stop -> Intro -> (one Intro frame) -> Fight  (BattleCore.cs:176-200, 262-291) */
static EnvironmentState run_round_start(Env *e, const fo_config *cfg) {
    e->roundState = RS_INTRO;
    f_SetupBattleStart(&e->fighter[0], -2.0f, 1);             /* :264 */
    f_SetupBattleStart(&e->fighter[1], 2.0f, 0);              /* :265 */
    /* :274-277.  P1's bot is NOT reset: the only way the reference runs a bot as P1 is by_example, which launches the game
     * with --p1-bot --p1-spectator (footsies.py:230-232); GameManager.cs:200-201 then wraps the bot actor in a
     * TrainingActorRemoteSpectator, and `trainingManager.actorP1 is TrainingBattleAIActor` (:274) is false for the wrapper.
     * P1's BattleAI therefore keeps its queues and its last recorded fight state across rounds, and its fightStates start
     * out null (first call of the process returns 0 without drawing).  Pinned by tests/test_oracle_vs_ref.py. */
    if (cfg->p2_bot) ai_Reset(e, &e->ai[1]);
    if (!cfg->stale_intro_input) { e->actorInput[0] = 0; e->actorInput[1] = 0; }
    UpdateIntroState(e);                                      /* next tick (:183-192), introStateTime = 0 (:125) */
    e->roundState = RS_FIGHT;                                 /* ChangeRoundState(Fight) :281-291 */
    e->frameCount = -1;
    e->currentRecordingInputIndex = 0;
    e->events = 0;
    EnvironmentState s = GetEnvironmentState(e);
    TrainingManager_Step(e, cfg, 0);
    return s;
}

/* One Fight-state FixedUpdate (BattleCore.cs:201-220) */
static EnvironmentState run_fight_tick(Env *e, const fo_config *cfg, int p1_action, int p2_action, int *battleOver) {
    if (!cfg->p1_bot) e->actorInput[0] = p1_action;           /* TrainingRemoteActor.cs:113-116 */
    if (!cfg->p2_bot) e->actorInput[1] = p2_action;
    e->frameCount++;
    e->events = 0;
    int prev_p1_action = e->fighter[0].currentActionID;
    UpdateFightState(e);
    *battleOver = f_isDead(&e->fighter[0]) || f_isDead(&e->fighter[1]);
    if (*battleOver) {                                        /* ChangeRoundState(KO) :292-305 */
        e->roundState = RS_KO;
        f_ClearInput(&e->fighter[0]);
        f_ClearInput(&e->fighter[1]);
    }
    EnvironmentState s = GetEnvironmentState(e);
    TrainingManager_Step(e, cfg, *battleOver);
    /* statistics (wrappers/statistics.py:26-50 semantics, evaluated every frame) */
    int a = e->fighter[0].currentActionID;
    if (a != prev_p1_action && (a == A_N_SPECIAL || a == A_B_SPECIAL)) {
        e->stats[FO_STAT_P1_SPECIALS]++;
        if (prev_p1_action != A_N_ATTACK && prev_p1_action != A_B_ATTACK) e->stats[FO_STAT_P1_SPECIALS_NEUTRAL]++;
    }
    e->frames_simulated++;
    e->episode_frames++;
    return s;
}

/* ------------------------------------------------------------------------------------------
 * Python side: FootsiesEnv (footsies-gym/footsies_gym/envs/footsies.py)
 * ---------------------------------------------------------------------------------------- */
/* FootsiesEnv._extract_obs (footsies.py:336-368) */
static void extract_obs(const EnvironmentState *s, float obs[8]) {
    int p1_simple = (s->p1Move == A_STAND || s->p1Move == A_FORWARD || s->p1Move == A_BACKWARD) ? 0 : s->p1MoveFrame;
    int p2_simple = (s->p2Move == A_STAND || s->p2Move == A_FORWARD || s->p2Move == A_BACKWARD) ? 0 : s->p2MoveFrame;
    obs[0] = (float)s->p1Guard; obs[1] = (float)s->p2Guard;
    obs[2] = (float)move_index(s->p1Move); obs[3] = (float)move_index(s->p2Move);
    obs[4] = (float)p1_simple; obs[5] = (float)p2_simple;
    obs[6] = s->p1Position; obs[7] = s->p2Position;
}
/* FootsiesEnv._get_sparse_reward (footsies.py:382-386) */
static double get_sparse_reward(const EnvironmentState *next, int terminated) {
    return terminated ? (next->p2Vital == 0 ? 1 : -1) : 0;
}
/* FootsiesEnv._get_dense_reward (footsies.py:388-405); Python floats are doubles */
static double get_dense_reward(Env *e, const EnvironmentState *state, const EnvironmentState *next, int terminated) {
    double reward = 0.0;
    if (next->p1Guard < state->p1Guard) reward -= 0.3;
    if (next->p2Guard < state->p2Guard) reward += 0.3;
    e->cumulative_episode_reward += reward;
    if (terminated) reward += (next->p2Vital == 0 ? 1 : -1) - e->cumulative_episode_reward;
    return reward;
}

static void fill_fighter_state(const Fighter *f, fo_fighter_state *o) {
    o->pos_x = f->pos_x; o->velocity_x = f->velocity_x;
    o->action_id = f->currentActionID; o->action_frame = f->currentActionFrame;
    o->hitstun = f->currentHitStunFrame; o->guard = f->guardHealth; o->vital = f->vitalHealth;
    o->hit_count = f->currentActionHitCount; o->buffer_id = f->bufferActionID; o->reserve_id = f->reserveDamageActionID;
    o->is_input_backward = f->isInputBackward; o->is_reserve_prox = f->isReserveProximityGuard;
    o->shake = f->spriteShakePosition; o->has_won = f->hasWon;
    o->input0 = f->input[0];
    o->hist_left = 0; o->hist_right = 0;
    for (int i = 0; i < 32; i++) {
        if (f->input[i] & IN_LEFT) o->hist_left |= 1u << i;
        if (f->input[i] & IN_RIGHT) o->hist_right |= 1u << i;
    }
    int run = 0;
    while (run < 59 && (f->input[run] & IN_ATTACK)) run++;
    o->attack_run = run;
}

static void fill_trace(Env *e, const EnvironmentState *obs_state, double reward, int terminated, int battleOver, int was_reset) {
    fo_trace *t = &e->last;
    fill_fighter_state(&e->fighter[0], &t->f[0]);
    fill_fighter_state(&e->fighter[1], &t->f[1]);
    t->frame = e->frameCount;
    t->recorded_input[0] = e->currentRecordingInputIndex > 0 ? e->lastRecordedInput[0] : 0;
    t->recorded_input[1] = e->currentRecordingInputIndex > 0 ? e->lastRecordedInput[1] : 0;
    t->events = e->events;
    t->battle_over = battleOver;
    t->was_reset = was_reset;
    t->rng_draws = e->rng_draws;
    memcpy(t->rng_state, e->rng, sizeof e->rng);
    t->bot_input[0] = e->actorInput[0]; t->bot_input[1] = e->actorInput[1];
    extract_obs(obs_state, t->obs);
    t->reward = (float)reward;
    t->reward_f64 = reward;
    t->terminated = terminated;
    t->info_frame = obs_state->globalFrame;                    /* _extract_info (footsies.py:370-380) */
    t->info_action[0] = obs_state->p1MostRecentAction; t->info_action[1] = obs_state->p2MostRecentAction;
    t->info_hitstun[0] = obs_state->p1Hitstun; t->info_hitstun[1] = obs_state->p2Hitstun;
}

/* FootsiesEnv.reset (footsies.py:482-515) */
static void env_reset(Env *e, const fo_config *cfg) {
    if (e->roundState == RS_KO) run_ko_end(e);                 /* game moved on by itself after a normal termination */
    /* else: RESET command -> ChangeRoundState(Stop) (BattleCore.cs:143-146), then Stop -> Intro in the same tick */
    EnvironmentState first = run_round_start(e, cfg);
    e->dq_len = 0;                                             /* delayed_frame_queue.clear() */
    e->cumulative_episode_reward = 0.0;
    e->current_state = first;
    while (e->dq_len < cfg->frame_delay) e->delayed[e->dq_len++] = first; /* maxlen - 1 copies (footsies.py:502-504) */
    e->has_terminated = 0;
    e->episode_frames = 0;
    e->episode_return = 0.0;
    fill_trace(e, &first, 0.0, 0, 0, 1);
}

/* FootsiesEnv.step (footsies.py:518-570), one frame */
static double env_step(Env *e, const fo_config *cfg, int p1_action, int p2_action) {
    EnvironmentState previous = e->current_state;
    int battleOver = 0;
    EnvironmentState most_recent = run_fight_tick(e, cfg, p1_action & 7, p2_action & 7, &battleOver);
    e->current_state = most_recent;
    e->delayed[e->dq_len++] = most_recent;                     /* append, then popleft (footsies.py:534-535) */
    EnvironmentState state = e->delayed[0];
    for (int i = 1; i < e->dq_len; i++) e->delayed[i - 1] = e->delayed[i];
    e->dq_len--;
    if (state.p1Move == A_DEAD || state.p1Move == A_WIN) state.p1Move = A_STAND; /* footsies.py:538-549 */
    if (state.p2Move == A_DEAD || state.p2Move == A_WIN) state.p2Move = A_STAND;
    int terminated = most_recent.p1Vital == 0 || most_recent.p2Vital == 0;       /* footsies.py:555 */
    double reward = cfg->dense_reward ? get_dense_reward(e, &previous, &most_recent, terminated)
                                      : get_sparse_reward(&most_recent, terminated);
    e->has_terminated = terminated;
    e->episode_return += reward;
    if (terminated) {
        e->stats[FO_STAT_EPISODES]++;
        if (most_recent.p1Vital == 0 && most_recent.p2Vital == 0) e->stats[FO_STAT_DOUBLE_KO]++;
        else if (most_recent.p2Vital == 0) e->stats[FO_STAT_P1_WINS]++;
        else e->stats[FO_STAT_P2_WINS]++;
        e->stats[FO_STAT_FRAMES] += e->episode_frames;
        e->return_sum += e->episode_return;
    }
    fill_trace(e, &state, reward, terminated, battleOver, 0);
    return reward;
}

/* ------------------------------------------------------------------------------------------
 * Batch API
 * ---------------------------------------------------------------------------------------- */
fo_batch *fo_create(int32_t num_envs, const fo_config *cfg, int64_t first_env_index) {
    fo_batch *b = (fo_batch *)calloc(1, sizeof *b);
    b->n = num_envs; b->cfg = *cfg; b->first_env_index = first_env_index;
    b->envs = (Env *)calloc((size_t)num_envs, sizeof(Env));
    for (int i = 0; i < num_envs; i++) {
        Env *e = &b->envs[i];
        e->ai[0].isPlayer1 = 1; e->ai[1].isPlayer1 = 0;
        e->fighter[0].maxSpriteShakeFrame = 6; e->fighter[1].maxSpriteShakeFrame = 6; /* Fighter.cs:110 */
        e->delayed = (EnvironmentState *)calloc((size_t)cfg->frame_delay + 2, sizeof(EnvironmentState));
        e->roundState = RS_STOP;
        e->has_terminated = 1;
        fo_rng_init(e->rng, (int32_t)(first_env_index + i));
    }
    return b;
}
void fo_destroy(fo_batch *b) {
    if (!b) return;
    for (int i = 0; i < b->n; i++) free(b->envs[i].delayed);
    free(b->envs); free(b);
}
void fo_seed(fo_batch *b, int64_t seed_base, const uint8_t *mask) {
    for (int i = 0; i < b->n; i++) if (!mask || mask[i]) {
        fo_rng_init(b->envs[i].rng, (int32_t)(seed_base + b->first_env_index + i));
        b->envs[i].tape = 0;
    }
}
void fo_set_rng_tape(fo_batch *b, int32_t env, const uint32_t *raw, int32_t n) {
    Env *e = &b->envs[env];
    uint32_t *copy = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(n > 0 ? n : 1)); /* leaked on purpose: test helper */
    memcpy(copy, raw, sizeof(uint32_t) * (size_t)n);
    e->tape = copy; e->tape_n = n; e->tape_pos = 0;
}
void fo_reset(fo_batch *b, const uint8_t *mask, fo_trace *out) {
    for (int i = 0; i < b->n; i++) {
        if (!mask || mask[i]) env_reset(&b->envs[i], &b->cfg);
        if (out) out[i] = b->envs[i].last;
    }
}
static void step_range(fo_batch *b, const uint8_t *actions_p1, const uint8_t *actions_p2, int repeat,
                       fo_trace *out, int lo, int hi) {
    const fo_config *cfg = &b->cfg;
    for (int i = lo; i < hi; i++) {
        Env *e = &b->envs[i];
        int a1 = actions_p1 ? actions_p1[i] : 0, a2 = actions_p2 ? actions_p2[i] : 0;
        if (e->has_terminated) {
            if (cfg->autoreset == 1) env_reset(e, cfg);         /* next-step autoreset: this call only resets */
            else { e->last.reward = 0.0f; e->last.reward_f64 = 0.0; e->last.was_reset = 0; } /* frozen */
        } else {
            double total = 0.0;
            for (int k = 0; k < repeat; k++) {
                total += env_step(e, cfg, a1, a2);
                if (e->has_terminated) break;
            }
            e->last.reward_f64 = total;
            e->last.reward = (float)total;
        }
        if (out) out[i] = e->last;
    }
}
typedef struct { fo_batch *b; const uint8_t *a1, *a2; int repeat; fo_trace *out; int lo, hi; } step_job;
static void *step_thread(void *p) {
    step_job *j = (step_job *)p;
    step_range(j->b, j->a1, j->a2, j->repeat, j->out, j->lo, j->hi);
    return 0;
}
void fo_step(fo_batch *b, const uint8_t *actions_p1, const uint8_t *actions_p2, int32_t repeat,
             fo_trace *out, int32_t num_threads) {
    if (repeat < 1) repeat = 1;
    if (num_threads > b->n) num_threads = b->n;
    if (num_threads <= 1) { step_range(b, actions_p1, actions_p2, repeat, out, 0, b->n); return; }
    if (num_threads > 256) num_threads = 256;
    pthread_t th[256]; step_job jobs[256];
    for (int t = 0; t < num_threads; t++) {   /* contiguous env blocks, one per thread */
        jobs[t] = (step_job){ b, actions_p1, actions_p2, repeat, out,
                              (int)((int64_t)b->n * t / num_threads), (int)((int64_t)b->n * (t + 1) / num_threads) };
        pthread_create(&th[t], 0, step_thread, &jobs[t]);
    }
    for (int t = 0; t < num_threads; t++) pthread_join(th[t], 0);
}
void fo_set_state(fo_batch *b, int32_t env, const fo_fighter_state *p1, const fo_fighter_state *p2, int32_t frame) {
    Env *e = &b->envs[env];
    const fo_fighter_state *src[2] = { p1, p2 };
    for (int i = 0; i < 2; i++) {
        Fighter *f = &e->fighter[i];
        const fo_fighter_state *s = src[i];
        f->pos_x = s->pos_x; f->pos_y = 0.0f; f->velocity_x = s->velocity_x; f->isFaceRight = (i == 0);
        f->currentActionID = s->action_id; f->currentActionFrame = s->action_frame;
        f->currentHitStunFrame = s->hitstun; f->guardHealth = s->guard; f->vitalHealth = s->vital;
        f->currentActionHitCount = s->hit_count; f->bufferActionID = s->buffer_id; f->reserveDamageActionID = s->reserve_id;
        f->isInputBackward = s->is_input_backward; f->isReserveProximityGuard = s->is_reserve_prox;
        f->spriteShakePosition = s->shake; f->hasWon = s->has_won;
        f_ClearInput(f);
        for (int k = INPUT_RECORD_FRAME - 1; k >= 0; k--) {   /* rebuild history oldest -> newest */
            int v = 0;
            if (k < 32 && (s->hist_left >> k & 1)) v |= IN_LEFT;
            if (k < 32 && (s->hist_right >> k & 1)) v |= IN_RIGHT;
            if (k < s->attack_run) v |= IN_ATTACK;
            f->input[k] = v;
        }
        for (int k = 0; k < INPUT_RECORD_FRAME; k++) {
            int prev = k + 1 < INPUT_RECORD_FRAME ? f->input[k + 1] : 0;
            f->inputDown[k] = (f->input[k] ^ prev) & f->input[k];
            f->inputUp[k] = (f->input[k] ^ prev) & ~f->input[k];
        }
        f_UpdateBoxes(f);
    }
    e->frameCount = frame;
    e->roundState = RS_FIGHT;
    e->has_terminated = 0;
    e->current_state = GetEnvironmentState(e);
    fill_trace(e, &e->current_state, 0.0, 0, 0, 0);
}
void fo_get_trace(fo_batch *b, int32_t env, fo_trace *out) { *out = b->envs[env].last; }

/* BattleCore.SaveState (BattleCore.cs:667-674) -> Fighter.SaveState (Fighter.cs:721-737) -> FighterState ctor
 * (FighterState.cs:58-131) */
void fo_save_battle_state(fo_batch *b, int32_t env, fo_battle_state *out) {
    const Env *e = &b->envs[env];
    memset(out, 0, sizeof *out);
    for (int i = 0; i < 2; i++) {
        const Fighter *f = &e->fighter[i];
        fo_full_fighter *o = &out->p[i];
        o->position[0] = f->pos_x; o->position[1] = f->pos_y;
        o->velocity_x = f->velocity_x;
        o->isFaceRight = f->isFaceRight;
        o->n_hitboxes = f->n_hitboxes;
        for (int k = 0; k < f->n_hitboxes; k++) {
            const Hitbox *h = &f->hitboxes[k];
            o->hitboxes[k].rect = (fo_rect){ h->rect.x, h->rect.y, h->rect.width, h->rect.height };
            o->hitboxes[k].proximity = h->proximity;
            o->hitboxes[k].attackID = h->attackID;
        }
        o->n_hurtboxes = f->n_hurtboxes;
        for (int k = 0; k < f->n_hurtboxes; k++) {
            const Rect *r = &f->hurtboxes[k].rect;
            o->hurtboxes[k] = (fo_rect){ r->x, r->y, r->width, r->height };
        }
        o->pushbox = (fo_rect){ f->pushbox.rect.x, f->pushbox.rect.y, f->pushbox.rect.width, f->pushbox.rect.height };
        o->vitalHealth = f->vitalHealth; o->guardHealth = f->guardHealth;
        o->currentActionID = f->currentActionID; o->currentActionFrame = f->currentActionFrame;
        o->currentActionHitCount = f->currentActionHitCount; o->currentHitStunFrame = f->currentHitStunFrame;
        for (int k = 0; k < INPUT_RECORD_FRAME; k++) {
            o->input[k] = f->input[k]; o->inputDown[k] = f->inputDown[k]; o->inputUp[k] = f->inputUp[k];
        }
        o->isInputBackward = f->isInputBackward; o->isReserveProximityGuard = f->isReserveProximityGuard;
        o->bufferActionID = f->bufferActionID; o->reserveDamageActionID = f->reserveDamageActionID;
        o->spriteShakePosition = f->spriteShakePosition; o->maxSpriteShakeFrame = f->maxSpriteShakeFrame;
        o->hasWon = f->hasWon;
    }
    out->roundStartTime = 0.0f;   /* Time.fixedTime at round start: display only */
    out->frameCount = e->frameCount;
}

/* BattleCore.LoadState (BattleCore.cs:677-683) -> Fighter.LoadState (Fighter.cs:739-811) */
void fo_load_battle_state(fo_batch *b, int32_t env, const fo_battle_state *in) {
    Env *e = &b->envs[env];
    for (int i = 0; i < 2; i++) {
        Fighter *f = &e->fighter[i];
        const fo_full_fighter *s = &in->p[i];
        f->pos_x = s->position[0]; f->pos_y = s->position[1];
        f->velocity_x = s->velocity_x;
        f->isFaceRight = s->isFaceRight;
        f->n_hitboxes = s->n_hitboxes;
        for (int k = 0; k < s->n_hitboxes; k++) {
            f->hitboxes[k].rect = (Rect){ s->hitboxes[k].rect.x, s->hitboxes[k].rect.y, s->hitboxes[k].rect.width, s->hitboxes[k].rect.height };
            f->hitboxes[k].attackID = s->hitboxes[k].attackID;
            f->hitboxes[k].proximity = s->hitboxes[k].proximity;
        }
        f->n_hurtboxes = s->n_hurtboxes;
        for (int k = 0; k < s->n_hurtboxes; k++)
            f->hurtboxes[k].rect = (Rect){ s->hurtboxes[k].x, s->hurtboxes[k].y, s->hurtboxes[k].width, s->hurtboxes[k].height };
        f->pushbox.rect = (Rect){ s->pushbox.x, s->pushbox.y, s->pushbox.width, s->pushbox.height };
        f->vitalHealth = s->vitalHealth; f->guardHealth = s->guardHealth;
        f->currentActionID = s->currentActionID; f->currentActionFrame = s->currentActionFrame;
        f->currentActionHitCount = s->currentActionHitCount; f->currentHitStunFrame = s->currentHitStunFrame;
        for (int k = 0; k < INPUT_RECORD_FRAME; k++) {
            f->input[k] = s->input[k]; f->inputDown[k] = s->inputDown[k]; f->inputUp[k] = s->inputUp[k];
        }
        f->isInputBackward = s->isInputBackward; f->isReserveProximityGuard = s->isReserveProximityGuard;
        f->bufferActionID = s->bufferActionID; f->reserveDamageActionID = s->reserveDamageActionID;
        f->spriteShakePosition = s->spriteShakePosition; f->maxSpriteShakeFrame = s->maxSpriteShakeFrame;
        f->hasWon = s->hasWon;
    }
    e->frameCount = in->frameCount;
    /* FootsiesEnv._current_state is NOT touched: the Python side only learns about the load from the next state it
     * receives, so its next dense reward compares guard bars across the load (footsies.py:530, 556-558).
     * Test convenience (not in the reference): refresh has_terminated and the trace so that fo_get_trace shows the loaded
     * fighters. */
    e->has_terminated = f_isDead(&e->fighter[0]) || f_isDead(&e->fighter[1]);
    fill_fighter_state(&e->fighter[0], &e->last.f[0]);
    fill_fighter_state(&e->fighter[1], &e->last.f[1]);
    e->last.frame = e->frameCount;
}
int64_t fo_frames_simulated(fo_batch *b) {
    int64_t s = 0;
    for (int i = 0; i < b->n; i++) s += b->envs[i].frames_simulated;
    return s;
}
void fo_stats(fo_batch *b, int64_t *out, double *return_sum) {
    double rs = 0.0;
    for (int k = 0; k < FO_STAT_COUNT; k++) out[k] = 0;
    for (int i = 0; i < b->n; i++) {
        for (int k = 0; k < FO_STAT_COUNT; k++) out[k] += b->envs[i].stats[k];
        rs += b->envs[i].return_sum;
    }
    if (return_sum) *return_sum = rs;
}
