"""Game-state records in the reference's own schema, and their mapping to the compact device state.

Mirrors footsies-gym/footsies_gym/state.py (FootsiesState :7-76, FootsiesBattleState :79-102,
FootsiesFighterState :105-137) -- same class names, field names, field order and JSON layout, which is the
layout `JsonUtility.ToJson` gives Assets/Script/BattleState.cs:9-24 and FighterState.cs:26-56 -- so that a
battle state saved by a real game build can be loaded into the GPU simulator and vice versa
(FootsiesEnv.save_battle_state / load_battle_state, footsies.py:432-444, remote-control STATE_SAVE / STATE_LOAD,
BattleCore.cs:667-683, Fighter.cs:721-811).

The device keeps 64 bytes per battle (csrc/state_codec.h), not the 3 x 180 ints of input history per fighter of
the C# class.  compact -> reference expands the history it has (Left/Right of the last 16 frames, the Attack
run length) into 180-entry arrays and rebuilds hit / hurt / push boxes from the frame data; reference -> compact
keeps exactly the part of the history the battle logic can still read (dash detection reads input[0..16],
Fighter.cs:585-635; the hold-release special reads Attack on input[1..59], :569-583) and drops the boxes (they
are rebuilt by UpdateBoxes before anything reads them, BattleCore.cs:347-364).  tests/test_battle_state.py
proves the round trip is behaviour-preserving against the oracle's full-history implementation.
"""
import dataclasses
import json
from typing import List

import numpy as np

from . import frame_data as _fd

INPUT_RECORD_FRAME = 180      # Fighter.cs:98 inputRecordFrame
_HIST_FRAMES = 16             # frames of Left/Right history kept on the device
_ATTACK_RUN_MAX = 59          # specialAttackHoldFrame - 1 (Fighter.cs:569-583)
MAX_SPRITE_SHAKE_FRAME = 6    # Fighter.cs:110 (a constant of the game; the kernel clamps with it)
_ACTIONS_BY_ID = {a["actionID"]: a for a in _fd.ACTIONS}
_f32 = np.float32


@dataclasses.dataclass
class FootsiesState:
    """The environment state the game reports every frame (EnvironmentState.cs:12-26)."""

    p1Vital: int
    p2Vital: int
    p1Guard: int
    p2Guard: int
    p1Move: int
    p2Move: int
    p1MoveFrame: int
    p2MoveFrame: int
    p1Position: float
    p2Position: float
    globalFrame: int
    p1MostRecentAction: "tuple[bool, bool, bool]"
    p2MostRecentAction: "tuple[bool, bool, bool]"
    p1Hitstun: int
    p2Hitstun: int

    def __post_init__(self):
        # the game sends the InputDefine bitmask; the reference unpacks it into (left, right, attack)
        if not isinstance(self.p1MostRecentAction, tuple):
            a = int(self.p1MostRecentAction)
            self.p1MostRecentAction = ((a & 1) != 0, (a & 2) != 0, (a & 4) != 0)
        if not isinstance(self.p2MostRecentAction, tuple):
            a = int(self.p2MostRecentAction)
            self.p2MostRecentAction = ((a & 1) != 0, (a & 2) != 0, (a & 4) != 0)

    @staticmethod
    def from_battle_state(battle_state: "FootsiesBattleState") -> "FootsiesState":
        p1, p2 = battle_state.p1State, battle_state.p2State
        return FootsiesState(
            p1Vital=p1.vitalHealth, p2Vital=p2.vitalHealth, p1Guard=p1.guardHealth, p2Guard=p2.guardHealth,
            p1Move=p1.currentActionID, p2Move=p2.currentActionID,
            p1MoveFrame=p1.currentActionFrame, p2MoveFrame=p2.currentActionFrame,
            p1Position=p1.position[0], p2Position=p2.position[0], globalFrame=battle_state.frameCount,
            p1MostRecentAction=p1.input[0], p2MostRecentAction=p2.input[0],
            p1Hitstun=p1.currentHitStunFrame, p2Hitstun=p2.currentHitStunFrame)


@dataclasses.dataclass
class FootsiesFighterState:
    """Full state of one fighter (FighterState.cs:26-56), field order as in the C# class."""

    position: List[float]
    velocity_x: float
    isFaceRight: bool

    hitboxes: List[dict]
    hurtboxes: List[dict]
    pushbox: dict

    vitalHealth: int
    guardHealth: int

    currentActionID: int
    currentActionFrame: int
    currentActionHitCount: int

    currentHitStunFrame: int

    input: List[int]
    inputDown: List[int]
    inputUp: List[int]

    isInputBackward: bool
    isReserveProximityGuard: bool

    bufferActionID: int
    reserveDamageActionID: int

    spriteShakePosition: int
    maxSpriteShakeFrame: int

    hasWon: bool


@dataclasses.dataclass
class FootsiesBattleState:
    """Full state of one battle (BattleState.cs:9-24)."""

    p1State: FootsiesFighterState
    p2State: FootsiesFighterState

    roundStartTime: float
    frameCount: int

    @staticmethod
    def from_json(battle_state_json: str) -> "FootsiesBattleState":
        d = json.loads(battle_state_json)
        return FootsiesBattleState(
            p1State=FootsiesFighterState(**d["p1State"]), p2State=FootsiesFighterState(**d["p2State"]),
            roundStartTime=d["roundStartTime"], frameCount=d["frameCount"])

    def json(self) -> str:
        return json.dumps(dataclasses.asdict(self))


# ---------------------------------------------------------------------------------------------------------
# boxes of an (action, frame) at a position: Fighter.ApplyCurrentActionData / TransformToFightRect
# (Fighter.cs:671-719) with ActionData.Get*Data (ActionData.cs:109-144); fp32 like the game
# ---------------------------------------------------------------------------------------------------------
def _fight_rect(rect, pos_x, face_right):
    x, y, w, h = (_f32(v) for v in rect)
    sign = _f32(1.0 if face_right else -1.0)
    return {"x": float(_f32(pos_x) + x * sign), "y": float(_f32(0.0) + y), "width": float(w), "height": float(h)}


def boxes_for(action_id: int, frame: int, pos_x: float, face_right: bool):
    """(hitboxes, hurtboxes, pushbox) in the FighterState JSON layout for the given action frame."""
    a = _ACTIONS_BY_ID[action_id]
    hit = [{"rect": _fight_rect(h["rect"], pos_x, face_right), "proximity": bool(h["proximity"]),
            "attackID": int(h["attackID"])}
           for h in a["hitboxes"] if h["se"][0] <= frame <= h["se"][1]]
    hurt = [_fight_rect(_fd.CONSTS["baseHurtBoxRect"] if h["useBaseRect"] else h["rect"], pos_x, face_right)
            for h in a["hurtboxes"] if h["se"][0] <= frame <= h["se"][1]]
    push = next((p for p in a["pushboxes"] if p["se"][0] <= frame <= p["se"][1]), None)   # first match (:135-144)
    prect = _fd.CONSTS["basePushBoxRect"] if push is None or push["useBaseRect"] else push["rect"]
    return hit, hurt, _fight_rect(prect, pos_x, face_right)


# ---------------------------------------------------------------------------------------------------------
# compact device state (one record of _capi.env_state_dtype()) <-> reference schema
# ---------------------------------------------------------------------------------------------------------
def _expand_inputs(hist_left: int, hist_right: int, attack_run: int):
    """input / inputDown / inputUp arrays (Fighter.cs:172-188) from the compact history; frames older than
    what the device keeps read as "nothing held"."""
    inp = [0] * INPUT_RECORD_FRAME
    for k in range(INPUT_RECORD_FRAME):
        v = 0
        if k < _HIST_FRAMES:
            v |= (hist_left >> k) & 1
            v |= ((hist_right >> k) & 1) << 1
        if k < attack_run:
            v |= 4
        inp[k] = v
    down = [0] * INPUT_RECORD_FRAME
    up = [0] * INPUT_RECORD_FRAME
    for k in range(INPUT_RECORD_FRAME):
        prev = inp[k + 1] if k + 1 < INPUT_RECORD_FRAME else 0
        down[k] = (inp[k] ^ prev) & inp[k]
        up[k] = (inp[k] ^ prev) & ~inp[k] & 7
    return inp, down, up


def fighter_to_reference(f, face_right: bool) -> FootsiesFighterState:
    """One `fg_fighter_state` record -> FootsiesFighterState."""
    inp, down, up = _expand_inputs(int(f["hist_left"]), int(f["hist_right"]), int(f["attack_run"]))
    action_id, frame, pos_x = int(f["action_id"]), int(f["action_frame"]), float(f["pos_x"])
    hit, hurt, push = boxes_for(action_id, frame, pos_x, face_right)
    shake = int(f["shake"])
    return FootsiesFighterState(
        position=[pos_x, 0.0], velocity_x=float(f["velocity_x"]), isFaceRight=bool(face_right),
        hitboxes=hit, hurtboxes=hurt, pushbox=push,
        vitalHealth=int(f["vital"]), guardHealth=int(f["guard"]),
        currentActionID=action_id, currentActionFrame=frame, currentActionHitCount=int(f["hit_count"]),
        currentHitStunFrame=int(f["hitstun"]), input=inp, inputDown=down, inputUp=up,
        isInputBackward=bool(f["is_input_backward"]), isReserveProximityGuard=bool(f["is_reserve_prox"]),
        bufferActionID=int(f["buffer_id"]), reserveDamageActionID=int(f["reserve_id"]),
        spriteShakePosition=shake, maxSpriteShakeFrame=MAX_SPRITE_SHAKE_FRAME, hasWon=bool(f["has_won"]))


def env_state_to_battle_state(rec) -> FootsiesBattleState:
    """One `fg_env_state` record (FootsiesEnv.get_state()[i]) -> FootsiesBattleState."""
    return FootsiesBattleState(
        p1State=fighter_to_reference(rec["f"][0], True), p2State=fighter_to_reference(rec["f"][1], False),
        roundStartTime=0.0,           # Time.fixedTime of the round start: display only (BattleCore.cs:283)
        frameCount=int(rec["frame"]))


class UnrepresentableStateError(ValueError):
    """The reference state cannot occur in a training battle of this simulator (e.g. hasWon / WIN action)."""


def fighter_from_reference(s: FootsiesFighterState, out):
    """FootsiesFighterState -> one `fg_fighter_state` record (filled in place)."""
    if len(s.input) < 1:
        raise UnrepresentableStateError("input history is empty")
    inp = list(s.input) + [0] * max(0, 60 - len(s.input))
    left = right = 0
    for k in range(_HIST_FRAMES):
        left |= (inp[k] & 1) << k
        right |= ((inp[k] >> 1) & 1) << k
    run = 0
    while run < _ATTACK_RUN_MAX and inp[run] & 4:
        run += 1
    if s.hasWon:
        raise UnrepresentableStateError("hasWon fighters only exist in the versus-mode End state")
    if s.bufferActionID not in (-1, 110) or s.reserveDamageActionID not in (-1, 310):
        raise UnrepresentableStateError("bufferActionID must be -1 or 110 and reserveDamageActionID -1 or 310")
    if s.maxSpriteShakeFrame != MAX_SPRITE_SHAKE_FRAME:
        raise UnrepresentableStateError("maxSpriteShakeFrame is the game constant 6 (Fighter.cs:110)")
    if s.currentActionID not in _ACTIONS_BY_ID:
        raise UnrepresentableStateError(f"unknown action id {s.currentActionID}")
    out["pos_x"] = _f32(s.position[0])
    out["velocity_x"] = _f32(s.velocity_x)
    out["action_id"] = s.currentActionID
    out["action_frame"] = s.currentActionFrame
    out["hitstun"] = s.currentHitStunFrame
    out["guard"] = s.guardHealth
    out["vital"] = s.vitalHealth
    out["hit_count"] = s.currentActionHitCount
    out["buffer_id"] = s.bufferActionID
    out["reserve_id"] = s.reserveDamageActionID
    out["is_input_backward"] = int(bool(s.isInputBackward))
    out["is_reserve_prox"] = int(bool(s.isReserveProximityGuard))
    out["shake"] = s.spriteShakePosition
    out["has_won"] = 0
    out["input0"] = inp[0]
    out["hist_left"] = left
    out["hist_right"] = right
    out["attack_run"] = run


def battle_state_into_env_state(state: FootsiesBattleState, rec):
    """Overwrite the fighters and the frame counter of one `fg_env_state` record with a reference battle state.
    Like BattleCore.LoadState (BattleCore.cs:677-683) nothing else changes: actors' held inputs, bot queues and
    the RNG stay as they are."""
    fighter_from_reference(state.p1State, rec["f"][0])
    fighter_from_reference(state.p2State, rec["f"][1])
    rec["frame"] = state.frameCount
    # the env reports input[0] of both fighters as p{1,2}MostRecentAction on the next state (BattleCore.cs:463-464
    # reads the recording buffer, which LoadState does not touch) -> keep rec["recorded_input"]
    rec["done"] = int(state.p1State.vitalHealth <= 0 or state.p2State.vitalHealth <= 0)
