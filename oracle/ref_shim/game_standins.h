/*
 * game_standins.h -- hand-written stand-ins for the reference classes that are NOT transliterated because they are
 * process plumbing (sockets, scene loading, audio, keyboard), i.e. boundary (ii) that the task deletes:
 *
 *   GameManager            (Assets/Script/GameManager.cs)            only the fields BattleCore reads; what Awake() builds
 *                                                                     from the command line is built by the harness
 *   TrainingRemoteControl  (Assets/Script/TrainingRemoteControl.cs)  commands are queued by the harness instead of being
 *                                                                     read from a socket; same Command enum and accessors
 *   InputManager           (Assets/Script/InputManager*.cs)          keyboard: nothing is ever pressed
 *   SoundManager           (Assets/Script/SoundManager.cs)           audio: no-op
 *
 * TEST INFRASTRUCTURE ONLY.  Included by the generated file between its forward declarations and its class definitions.
 */
#pragma once
#include "unity_shim.h"

namespace Footsies {

struct TrainingRemoteControl : Object {
    enum class Command : int { NONE = 0, RESET = 1, STATE_SAVE = 2, STATE_LOAD = 3, P2_BOT = 4, SEED = 5 };   /* TrainingRemoteControl.cs:18-26 */
    Ref<BattleState> battleState;         /* STATE_LOAD payload */
    Ref<BattleState> savedState;          /* what SendBattleState "sent" */
    Ref<TrainingActor> p2Saved;
    Ref<TrainingBattleAIActor> p2Bot;
    bool isP2Bot = false;
    int seed = 0;
    Command pending = Command::NONE;      /* set by the harness; consumed by the next FixedUpdate */
    Command ProcessCommand() { Command c = pending; pending = Command::NONE; return c; }   /* :81-108 minus the socket */
    Ref<BattleState> GetDesiredBattleState() { return battleState; }
    void SetP2Saved(Ref<TrainingActor> p2) { p2Saved = p2; }
    void SetP2Bot(Ref<TrainingBattleAIActor> p2) { p2Bot = p2; }
    void SendBattleState(Ref<BattleState> state) { savedState = state; }
    void Close() {}
};

struct GameManager : Object {
    static thread_local GameManager *Instance;     /* Singleton<GameManager>.Instance; bound by the harness per game */
    bool isVsCPU = true;                  /* training always loads the vs-CPU scene (GameManager.cs:214-217, 245-249) */
    Ref<TrainingManager> trainingManager;
    Ref<TrainingRemoteControl> trainingRemoteControl;
    Ref<TrainingBattleAIActor> botP1, botP2;
    void LoadTitleScene() {}
};

struct InputButton { bool IsPressed() { return false; } bool WasPressedThisFrame() { return false; }
                     InputButton *operator->() { return this; } };
struct GameplayActions { InputButton p1Left, p1Right, p1Attack, p2Left, p2Right, p2Attack, debugPause, debugPauseAdvance, cancel;
                         GameplayActions *operator->() { return this; } };
struct InputManager : Object {
    static InputManager *Instance;
    GameplayActions gameplay;
};
struct SoundManager : Object {
    static SoundManager *Instance;
    void playFighterSE(Ref<AudioClip>, bool, float) {}
};

}  // namespace Footsies
