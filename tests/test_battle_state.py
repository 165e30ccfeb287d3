"""save_battle_state / load_battle_state in the reference's JSON schema.

Reference: FootsiesEnv.save_battle_state / load_battle_state (footsies.py:432-444), state.py:78-137,
BattleCore.SaveState / LoadState (BattleCore.cs:667-683), Fighter.SaveState / LoadState (Fighter.cs:721-811).

CPU tests pin (a) the JSON schema to a fixture that went through the reference's own dataclasses
(tests/golden/make_golden.py) and (b) the compact <-> full-history mapping of footsies_gym_b200/state.py: an oracle
battle with the full 180-entry histories is saved, squeezed through the 64-byte device representation and loaded
into a second oracle, and both must stay identical frame by frame.  GPU tests do the same across the C ABI.
"""
import json
import os

import numpy as np
import pytest

from footsies_gym_b200 import _capi
from parity import canonical_history
from footsies_gym_b200.state import (FootsiesBattleState, FootsiesState, UnrepresentableStateError,
                                     battle_state_into_env_state, env_state_to_battle_state)

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ref_battle_state.json")
L, R, A = 1, 2, 4


def sticky_tape(rng, steps, n, p_change=0.2):
    """Inputs that are held for a while, so that dashes, charged specials and blocks all happen."""
    cur = rng.integers(0, 8, size=n, dtype=np.uint8)
    out = np.zeros((steps, n), dtype=np.uint8)
    for t in range(steps):
        change = rng.random(n) < p_change
        cur = np.where(change, rng.integers(0, 8, size=n, dtype=np.uint8), cur).astype(np.uint8)
        out[t] = cur
    return out


def squeeze_through_compact(state_dict):
    """reference JSON dict -> FootsiesBattleState -> 64-byte device record fields -> FootsiesBattleState -> dict."""
    bs = FootsiesBattleState.from_json(json.dumps(state_dict))
    rec = np.zeros(1, dtype=_capi.env_state_dtype())
    battle_state_into_env_state(bs, rec[0])
    return json.loads(env_state_to_battle_state(rec[0]).json())


def test_golden_schema_round_trip():
    """The fixture was emitted by the reference's own FootsiesBattleState.json(); our classes must parse it, expose
    the same FootsiesState view and re-emit it byte for byte."""
    fx = json.load(open(GOLDEN))
    for case in fx["cases"]:
        bs = FootsiesBattleState.from_json(case["reference_json"])
        assert bs.json() == case["reference_json"]
        st = FootsiesState.from_battle_state(bs)
        exp = case["reference_footsies_state"]
        for k, v in exp.items():
            got = getattr(st, k)
            assert (list(got) if isinstance(got, tuple) else got) == v, (k, got, v)
    assert fx["fighter_fields"] == [f.name for f in __import__("dataclasses").fields(type(bs.p1State))]


def test_compact_mapping_preserves_behaviour(oracle):
    """Oracle A (full histories) -> save -> 64-byte representation -> load into oracle B -> lockstep."""
    n, stop, follow = 96, 450, 400
    rng = np.random.default_rng(2024)
    a = oracle.OracleBatch(n, p2_bot=False, seed=0)
    b = oracle.OracleBatch(n, p2_bot=False, seed=0)
    a.reset()
    b.reset()
    t1, t2 = sticky_tape(rng, stop + follow, n), sticky_tape(rng, stop + follow, n)
    for t in range(stop):
        a.step(t1[t], t2[t])
        b.step(t2[t], t1[t])                         # B lives a different life until the transplant
    # battles that are over on either side are between rounds (LoadState does not change the round state)
    alive = ~(a.trace["terminated"].astype(bool) | b.trace["terminated"].astype(bool))
    alive &= ~(a.trace["was_reset"].astype(bool) | b.trace["was_reset"].astype(bool))
    kept = ("position", "velocity_x", "vitalHealth", "guardHealth", "currentActionID", "currentActionFrame",
            "currentActionHitCount", "currentHitStunFrame", "isInputBackward", "isReserveProximityGuard",
            "bufferActionID", "reserveDamageActionID", "spriteShakePosition", "maxSpriteShakeFrame", "hasWon",
            "isFaceRight")
    for i in np.nonzero(alive)[0]:
        full = a.save_battle_state(i)
        small = squeeze_through_compact(full)
        for side in ("p1State", "p2State"):
            for key in kept:
                assert small[side][key] == full[side][key], (side, key)
            # of the Left/Right history the device keeps the canonical form dash detection can read (tests/parity.py)
            fl = sum((v & 1) << k for k, v in enumerate(full[side]["input"][:16]))
            fr = sum(((v >> 1) & 1) << k for k, v in enumerate(full[side]["input"][:16]))
            sl = sum((v & 1) << k for k, v in enumerate(small[side]["input"][:16]))
            sr = sum(((v >> 1) & 1) << k for k, v in enumerate(small[side]["input"][:16]))
            cf, cs = canonical_history(np.array([fl]), np.array([fr])), canonical_history(np.array([sl]), np.array([sr]))
            assert int(cf[0][0]) == int(cs[0][0]) and int(cf[1][0]) == int(cs[1][0])
        assert small["frameCount"] == full["frameCount"]
        b.load_battle_state(i, small)
    assert alive.sum() >= n // 3
    fields = ("pos_x", "velocity_x", "action_id", "action_frame", "hitstun", "guard", "vital", "hit_count",
              "buffer_id", "reserve_id", "is_input_backward", "is_reserve_prox", "shake", "input0", "attack_run")
    compared = 0
    for t in range(stop, stop + follow):
        a.step(t1[t], t2[t])
        b.step(t1[t], t2[t])
        # the round restart replays the actors' stale inputs, which are not part of a battle state: follow every
        # battle until it ends
        alive &= ~(a.trace["was_reset"].astype(bool) | b.trace["was_reset"].astype(bool))
        ta, tb = a.trace[alive], b.trace[alive]
        for f in fields:
            assert np.array_equal(ta["f"][f], tb["f"][f]), f"frame {t}: {f}"
        for f in ("obs", "frame", "terminated", "info_action", "info_hitstun"):
            assert np.array_equal(ta[f], tb[f]), f"frame {t}: {f}"
        # the terminal dense reward compensates the episode's accumulated reward, which is Python-side state and
        # not part of a battle state (footsies.py:399-403): compare the per-step part only
        # ... and not on the first step after the load, where the reference compares the guard bars with those of the last
        # state b's agent had received before the load (footsies.py:530, 556-558)
        live = ~ta["terminated"].astype(bool)
        if t > stop:
            assert np.array_equal(ta["reward"][live], tb["reward"][live]), f"frame {t}: reward"
        compared += int(alive.sum())
    assert compared > 3000


def test_long_attack_hold_survives_the_compact_form(oracle):
    """A special charged for 59+ frames before the save must still fire after the load (the run length is all the
    device keeps of the Attack history)."""
    a = oracle.OracleBatch(1, p2_bot=False, seed=0)
    b = oracle.OracleBatch(1, p2_bot=False, seed=0)
    a.reset()
    b.reset()
    for _ in range(70):
        a.step(np.array([A], np.uint8), np.array([0], np.uint8))
    b.load_battle_state(0, squeeze_through_compact(a.save_battle_state(0)))
    for x in (a, b):
        x.step(np.array([0], np.uint8), np.array([0], np.uint8))     # release -> N_SPECIAL (110)
    assert a.trace["f"]["action_id"][0, 0] == 110
    assert b.trace["f"]["action_id"][0, 0] == 110


def test_dash_history_survives_the_compact_form(oracle):
    a = oracle.OracleBatch(1, p2_bot=False, seed=0)
    b = oracle.OracleBatch(1, p2_bot=False, seed=0)
    a.reset()
    b.reset()
    z = np.array([0], np.uint8)
    for v in (0, 0, R, 0):                                            # tap forward, release ...
        a.step(np.array([v], np.uint8), z)
    b.load_battle_state(0, squeeze_through_compact(a.save_battle_state(0)))
    for x in (a, b):
        x.step(np.array([R], np.uint8), z)                            # ... tap again -> DASH_FORWARD (10)
    assert a.trace["f"]["action_id"][0, 0] == 10
    assert b.trace["f"]["action_id"][0, 0] == 10


def test_unrepresentable_states_are_rejected(oracle):
    a = oracle.OracleBatch(1, p2_bot=False, seed=0)
    a.reset()
    d = a.save_battle_state(0)
    rec = np.zeros(1, dtype=_capi.env_state_dtype())
    for key, val in (("hasWon", True), ("bufferActionID", 100), ("currentActionID", 999), ("maxSpriteShakeFrame", 3)):
        bad = json.loads(json.dumps(d))
        bad["p1State"][key] = val
        with pytest.raises(UnrepresentableStateError):
            battle_state_into_env_state(FootsiesBattleState.from_json(json.dumps(bad)), rec[0])


def test_saved_boxes_match_the_oracle(oracle):
    """Boxes rebuilt from the frame data equal the boxes the oracle's Fighter.UpdateBoxes built (up to the rounding of
    the push / wall displacement that ApplyPositionChange added to them afterwards), except on frames where a hit
    replaced the action after the boxes were built -- there the game's boxes are one frame stale by design."""
    n = 32
    rng = np.random.default_rng(5)
    a = oracle.OracleBatch(n, p2_bot=False, seed=0)
    a.reset()
    t1, t2 = sticky_tape(rng, 300, n), sticky_tape(rng, 300, n)
    checked = 0
    for t in range(300):
        a.step(t1[t], t2[t])
        if t % 7:
            continue
        for i in range(n):
            if a.trace["events"][i] != 0 or a.trace["was_reset"][i] or a.trace["terminated"][i]:
                continue    # a hit changed the action after the boxes were built: the game's boxes are one frame stale
            full = a.save_battle_state(i)
            small = squeeze_through_compact(full)
            for side in ("p1State", "p2State"):
                f, s = full[side], small[side]
                assert len(f["hitboxes"]) == len(s["hitboxes"]) and len(f["hurtboxes"]) == len(s["hurtboxes"])
                for hb_f, hb_s in zip(f["hitboxes"], s["hitboxes"]):
                    assert hb_f["attackID"] == hb_s["attackID"] and hb_f["proximity"] == hb_s["proximity"]
                    assert hb_f["rect"]["width"] == hb_s["rect"]["width"] and hb_f["rect"]["y"] == hb_s["rect"]["y"]
                    assert abs(hb_f["rect"]["x"] - hb_s["rect"]["x"]) < 1e-5
                for hu_f, hu_s in zip(f["hurtboxes"], s["hurtboxes"]):
                    assert hu_f["width"] == hu_s["width"] and hu_f["height"] == hu_s["height"]
                    assert abs(hu_f["x"] - hu_s["x"]) < 1e-5
                assert f["pushbox"]["width"] == s["pushbox"]["width"]
                checked += 1
    assert checked > 1000


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_gpu_save_then_oracle_follows(oracle):
    import torch
    from footsies_gym_b200 import FootsiesEnv
    n, warm, follow = 128, 350, 300
    rng = np.random.default_rng(77)
    env = FootsiesEnv(num_envs=n, device="cuda:0", opponent="self_play", seed=0)
    orc = oracle.OracleBatch(n, p2_bot=False, seed=0)
    env.reset()
    orc.reset()
    t1, t2 = sticky_tape(rng, warm + follow, n), sticky_tape(rng, warm + follow, n)
    for t in range(warm):
        env.step(torch.from_numpy(t1[t]), torch.from_numpy(t2[t]))
        orc.step(t2[t], t1[t])                       # the oracle lives a different life
    states = env.save_battle_state()
    assert len(states) == n and isinstance(states[0], FootsiesBattleState)
    alive = ~env.terminated.cpu().numpy()
    for i in range(n):
        orc.load_battle_state(i, json.loads(states[i].json()))
    for t in range(warm, warm + follow):
        env.step(torch.from_numpy(t1[t]), torch.from_numpy(t2[t]))
        orc.step(t1[t], t2[t])
        ks = env.get_state()
        alive &= ks["frame"] != -1
        alive &= ~orc.trace["was_reset"].astype(bool)
        for f in ("pos_x", "velocity_x", "action_id", "action_frame", "hitstun", "guard", "vital", "hit_count",
                  "buffer_id", "reserve_id", "is_input_backward", "is_reserve_prox", "shake", "attack_run"):
            assert np.array_equal(ks["f"][f][alive], orc.trace["f"][f][alive]), f"frame {t}: {f}"
        assert np.array_equal(env.obs.cpu().numpy()[alive], orc.trace["obs"][alive])
    assert alive.sum() > 0


@pytest.mark.gpu
def test_oracle_save_then_gpu_follows(oracle):
    import torch
    from footsies_gym_b200 import FootsiesEnv
    n, warm, follow = 128, 350, 300
    rng = np.random.default_rng(78)
    env = FootsiesEnv(num_envs=n, device="cuda:0", opponent="self_play", seed=0)
    orc = oracle.OracleBatch(n, p2_bot=False, seed=0)
    env.reset()
    orc.reset()
    t1, t2 = sticky_tape(rng, warm + follow, n), sticky_tape(rng, warm + follow, n)
    for t in range(warm):
        env.step(torch.from_numpy(t2[t]), torch.from_numpy(t1[t]))
        orc.step(t1[t], t2[t])
    alive = ~orc.trace["terminated"].astype(bool)
    for i in range(n):
        env.load_battle_state(json.dumps(orc.save_battle_state(i)), i)     # the full 180-entry JSON of the game
    ks = env.get_state()
    assert np.array_equal(ks["f"]["action_id"], orc.trace["f"]["action_id"])
    for t in range(warm, warm + follow):
        env.step(torch.from_numpy(t1[t]), torch.from_numpy(t2[t]))
        orc.step(t1[t], t2[t])
        ks = env.get_state()
        alive &= ks["frame"] != -1
        alive &= ~orc.trace["was_reset"].astype(bool)
        for f in ("pos_x", "velocity_x", "action_id", "action_frame", "hitstun", "guard", "vital", "hit_count",
                  "buffer_id", "reserve_id", "is_input_backward", "is_reserve_prox", "shake", "attack_run"):
            assert np.array_equal(ks["f"][f][alive], orc.trace["f"][f][alive]), f"frame {t}: {f}"
        assert np.array_equal(env.obs.cpu().numpy()[alive], orc.trace["obs"][alive])
        live = alive & ~orc.trace["terminated"].astype(bool)
        if t > warm:          # the step after the load is rewarded against the guard bars the env's agent saw before it
            assert np.array_equal(env.reward.cpu().numpy()[live], orc.trace["reward"][live])
    assert alive.sum() > 0


@pytest.mark.gpu
def test_dense_reward_across_a_load_follows_the_reference(oracle):
    """footsies.py:530, 556-558: the reward of the step after load_battle_state compares the new guard bars with those of
    the last state received BEFORE the load, and the cumulative reward of the episode runs on.  GPU env and oracle live the
    same life, both load the same earlier state in mid-episode; every reward afterwards (terminal ones included) agrees."""
    import torch
    from footsies_gym_b200 import FootsiesEnv
    n, steps = 96, 700
    rng = np.random.default_rng(79)
    env = FootsiesEnv(num_envs=n, device="cuda:0", opponent="self_play", seed=0, autoreset=False)
    orc = oracle.OracleBatch(n, p2_bot=False, seed=0, autoreset=False)
    env.reset()
    orc.reset()
    t1 = sticky_tape(rng, steps, n)
    t2 = sticky_tape(rng, steps, n)
    saved, loads, differing = {}, 0, 0
    for t in range(steps):
        env.step(torch.from_numpy(t1[t]), torch.from_numpy(t2[t]))
        orc.step(t1[t], t2[t])
        assert np.array_equal(env.reward.cpu().numpy(), orc.trace["reward"]), f"step {t}"
        assert np.array_equal(env.terminated.cpu().numpy().astype(np.int32), orc.trace["terminated"]), f"step {t}"
        assert np.array_equal(env.get_state()["f"]["guard"], orc.trace["f"]["guard"]), f"step {t}"
        live = ~orc.trace["terminated"].astype(bool)
        if t % 45 == 10:                              # remember an early state of every running battle ...
            for i in np.flatnonzero(live):
                saved.setdefault(int(i), orc.save_battle_state(int(i)))
        if t % 45 == 40:                              # ... and throw the battle back to it later in the same episode
            for i in np.flatnonzero(live):
                st = saved.pop(int(i), None)
                if st is None:
                    continue
                now = orc.trace["f"]["guard"][i]
                differing += int(st["p1State"]["guardHealth"] != now[0] or st["p2State"]["guardHealth"] != now[1])
                env.load_battle_state(json.dumps(st), int(i))
                orc.load_battle_state(int(i), st)
                loads += 1
    assert loads > 50 and differing > 5 and int(orc.trace["terminated"].sum()) > 10
