"""Known-answer tests that pin the CPU oracle to the C# text.

The reference ships no tests; these answers are derived by hand from the cited C# lines and the
F00 frame data, then checked against the oracle.  Self-play mode (both players from tapes), no RNG,
unless stated.  L=1, R=2, A=4.  P1 faces right (forward = R), P2 faces left (forward = L).
"""
import numpy as np
import pytest

L, R, A = 1, 2, 4
f32 = np.float32
DT = f32(0.02)


@pytest.fixture(params=["oracle", "reference_transliteration"])
def oracle(request):
    """Every known answer is asked of BOTH checkers: the hand-written restatement (oracle/) and the reference's own C#
    transliterated by tools/cs2cpp.py (oracle/_ref/) -- a known answer derived from a misreading of the C# fails the second."""
    import oracle_binding
    if request.param == "oracle":
        oracle_binding.build()
        return oracle_binding
    import types
    import ref_binding
    if not ref_binding.available():
        pytest.skip("oracle/_ref not built and /root/reference not present")
    ref_binding.build()
    return types.SimpleNamespace(OracleBatch=ref_binding.RefBatch)


def make(oracle, n=1, **kw):
    kw.setdefault("p2_bot", False)
    kw.setdefault("autoreset", False)
    b = oracle.OracleBatch(n, **kw)
    b.reset()
    return b


def run(b, tape1, tape2=None):
    """Step env 0 through the tapes, return list of copies of the trace."""
    out = []
    for i, a in enumerate(tape1):
        a2 = 0 if tape2 is None else tape2[i]
        out.append(b.step([a], [a2]).copy()[0])
    return out


def place(b, x1, x2, **kw):
    p1 = {"pos_x": x1, "action_frame": 1}
    p2 = {"pos_x": x2, "action_frame": 1}
    p1.update(kw.get("p1", {}))
    p2.update(kw.get("p2", {}))
    b.set_state(0, p1, p2, frame=0)


def test_reset_state_is_frame_minus_one_with_one_intro_frame(oracle):
    # App. B-2: one Intro frame runs before Fight -> MoveFrame 1, globalFrame -1 (BattleCore.cs:183-192, 281-291)
    b = make(oracle)
    t = b.trace[0]
    assert t["frame"] == -1 and t["info_frame"] == -1
    assert t["f"]["action_id"].tolist() == [0, 0] and t["f"]["action_frame"].tolist() == [1, 1]
    assert t["f"]["pos_x"].tolist() == [-2.0, 2.0]
    assert t["f"]["guard"].tolist() == [3, 3] and t["f"]["vital"].tolist() == [1, 1]
    assert t["obs"].tolist() == [3, 3, 0, 0, 0, 0, -2.0, 2.0]
    assert t["info_action"].tolist() == [0, 0] and t["terminated"] == 0 and t["reward"] == 0


def test_c1_idle_stand_period_24(oracle):
    # Fighter.cs:156-165, 474-478: STAND restarts when frame reaches frameCount 24
    b = make(oracle)
    tr = run(b, [0] * 50)
    frames = [t["f"]["action_frame"][0] for t in tr]
    exp, f = [], 1
    for _ in range(50):
        f += 1
        if f >= 24:
            f = 0
        exp.append(f)
    assert frames == exp
    assert all(t["obs"][4] == 0 and t["obs"][5] == 0 for t in tr)
    assert all(t["f"]["pos_x"].tolist() == [-2.0, 2.0] for t in tr)
    assert [t["frame"] for t in tr] == list(range(50))
    assert not any(t["terminated"] for t in tr)


def test_c2_walk_and_wall_clamp(oracle):
    # Fighter.cs:297-307: x += 2.2*sign*dt (P1 fwd); P2 holding R walks backward: x -= 1.8*(-1)*dt
    b = make(oracle)
    tr = run(b, [R] * 300, [R] * 300)
    x1, x2 = f32(-2.0), f32(2.0)
    fwd = f32(f32(2.2) * f32(1)) * DT
    bwd = f32(f32(1.8) * f32(-1)) * DT
    assert tr[0]["f"]["action_id"].tolist() == [1, 2] and tr[0]["f"]["action_frame"].tolist() == [0, 0]
    for t in tr:
        x1 = f32(x1 + fwd)
        x2 = f32(x2 - bwd)
        # wall clamp for base pushbox (w=1.4): |x| <= 4.3 (BattleCore.cs:503-519)
        xmax = f32(x2 + f32(f32(1.4) / f32(2)))
        if xmax > f32(5):
            x2 = f32(x2 + f32(f32(5) - xmax))
        # push when P1 reaches P2 (Rect semantics): handled below by breaking out early
        if f32(x2 - x1) < f32(1.4):
            break
        assert t["f"]["pos_x"][0] == x1 and t["f"]["pos_x"][1] == x2, t["frame"]
    assert abs(float(tr[-1]["f"]["pos_x"][1]) - 4.3) < 1e-5


def test_c3_push_keeps_distance_1_4(oracle):
    # BattleCore.cs:483-501: both displaced by half the overlap each frame
    b = make(oracle)
    tr = run(b, [R] * 200, [0] * 200)
    d = [float(t["f"]["pos_x"][1]) - float(t["f"]["pos_x"][0]) for t in tr]
    first = next(i for i, v in enumerate(d) if v < 1.4 + 1e-6)
    assert first > 40
    corner = next(i for i, t in enumerate(tr) if t["f"]["pos_x"][1] > 4.29)
    assert all(abs(v - 1.4) < 1e-5 for v in d[first:corner])
    # P2 (idle) is pushed back until the wall; there push-then-clamp (no iteration) leaves them overlapped (App. B-8)
    assert abs(float(tr[-1]["f"]["pos_x"][1]) - 4.3) < 1e-5 and d[-1] < 1.4 - 0.04
    assert tr[-1]["f"]["pos_x"].tolist() == tr[-2]["f"]["pos_x"].tolist()


@pytest.mark.parametrize("tape", [[R, 0, R], [R, R, 0, R]])
def test_c4_forward_dash(oracle, tape):
    # Fighter.cs:585-609 with dashAllowFrame 9; DASH_FORWARD.asset movements
    b = make(oracle)
    tr = run(b, tape + [0] * 20)
    k = len(tape) - 1
    assert tr[k]["f"]["action_id"][0] == 10 and tr[k]["f"]["action_frame"][0] == 0
    vel = [5] * 3 + [7] * 6 + [5] * 3 + [2] * 2 + [1] + [0]
    x = f32(tr[k - 1]["f"]["pos_x"][0])
    for i, v in enumerate(vel):
        x = f32(x + f32(f32(v) * f32(1)) * DT) if v else x
        t = tr[k + i]
        assert t["f"]["action_id"][0] == 10 and t["f"]["action_frame"][0] == i
        assert t["f"]["pos_x"][0] == x
        assert t["f"]["velocity_x"][0] == v
    assert tr[k + 16]["f"]["action_id"][0] == 0 and tr[k + 16]["f"]["action_frame"][0] == 0


def test_dash_rejected_cases(oracle):
    # opposite direction inside the window (Fighter.cs:592-595), and no neutral gap
    for tape in ([R, L, R], [R, R, R], [R] + [0] * 9 + [R]):
        b = make(oracle)
        tr = run(b, tape + [0] * 3)
        assert all(t["f"]["action_id"][0] != 10 for t in tr), tape
    # 7 neutral frames between taps is still inside dashAllowFrame 9 (i = 8)
    b = make(oracle)
    tr = run(b, [R] + [0] * 7 + [R])
    assert tr[-1]["f"]["action_id"][0] == 10


def test_backward_dash_and_invincibility_window(oracle):
    # P1 back = L.  DASH_BACKWARD.asset: frames 0-3 have no hurtbox (App. A)
    b = make(oracle)
    tr = run(b, [L, 0, L] + [0] * 25)
    assert tr[2]["f"]["action_id"][0] == 11 and tr[2]["f"]["action_frame"][0] == 0
    vel = [-10] * 3 + [-5] * 6 + [-3] * 4 + [-1] * 2 + [0]
    x = f32(tr[1]["f"]["pos_x"][0])
    for i, v in enumerate(vel):
        x = f32(x + f32(f32(v) * f32(1)) * DT) if v else x
        # wall clamp with the 0.8 wide dash pushbox: x >= -4.6
        xmin = f32(x - f32(f32(0.8) / f32(2)))
        if xmin < f32(-5):
            x = f32(x + f32(f32(-5) - xmin))
        assert tr[2 + i]["f"]["pos_x"][0] == x, i
    assert tr[2 + 21]["f"]["action_id"][0] == 11 and tr[2 + 22]["f"]["action_id"][0] == 0


def test_c5_whiffed_n_attack(oracle):
    b = make(oracle)
    tr = run(b, [A] + [0] * 30)
    for i in range(22):
        assert tr[i]["f"]["action_id"][0] == 100 and tr[i]["f"]["action_frame"][0] == i
    assert tr[22]["f"]["action_id"][0] == 0 and tr[22]["f"]["action_frame"][0] == 0
    assert all(t["events"] == 0 for t in tr)
    assert all(t["f"]["guard"].tolist() == [3, 3] for t in tr)
    # obs move index of N_ATTACK is 5 and move_frame is the raw frame (footsies.py:339-348)
    assert tr[3]["obs"][2] == 5 and tr[3]["obs"][4] == 3


def test_c6_hit_and_stun(oracle):
    # N_ATTACK real hitbox frames 4-5, reach x+0.9+0.9 = 1.8; base hurtbox half width 0.75 -> hits iff distance <= 2.55
    b = make(oracle)
    place(b, -1.0, 1.5)                          # distance 2.5
    tr = run(b, [A] + [0] * 40)
    assert all(t["events"] & 3 == 0 for t in tr[:4])
    t = tr[4]
    assert t["events"] & 1 and (t["events"] >> 4) & 3 == 1    # P1 connected, P2 result Damage
    assert t["f"]["action_id"].tolist() == [100, 200] and t["f"]["action_frame"].tolist() == [4, 0]
    assert t["f"]["guard"].tolist() == [3, 2] and t["f"]["hitstun"].tolist() == [12, 12]
    assert t["f"]["hit_count"].tolist() == [1, 0]
    assert t["f"]["shake"][1] == 4               # 12/3, P2 faces left -> +4 (Fighter.cs:438-444)
    assert t["reward"] == f32(0.3)
    # frozen for 12 frames (App. B-4), frame counter resumes on t+13
    for i in range(1, 13):
        assert tr[4 + i]["f"]["action_frame"].tolist() == [4, 0], i
        assert tr[4 + i]["f"]["hitstun"].tolist() == [12 - i, 12 - i]
    assert tr[4 + 13]["f"]["action_frame"].tolist() == [5, 1]
    # hit again impossible: hitCount stays 1 while frame 5 hitbox is active
    assert tr[4 + 13]["events"] & 1 == 0
    # out of range: distance 2.56 does not connect
    b = make(oracle)
    place(b, -1.0, 1.5625)
    tr = run(b, [A] + [0] * 10)
    assert all(t["events"] & 3 == 0 for t in tr)


def test_c7_block_then_guard_break(oracle):
    # P2 holds back (R) -> BACKWARD -> blocks N_ATTACK with GUARD_CROUCH(306); 4th block at guard 0 -> break
    b = make(oracle)
    place(b, 2.4, 4.3)                           # P2 cornered so that holding back does not walk out of range
    tape1, tape2 = [], []
    for _ in range(4):
        tape1 += [A] + [0] * 40
        tape2 += [R] * 41
    tape1 += [0] * 60
    tape2 += [R] * 60
    tr = run(b, tape1, tape2)
    blocks = [t for t in tr if t["events"] & 1]
    assert len(blocks) == 4
    for i, t in enumerate(blocks[:3]):
        assert (t["events"] >> 4) & 3 == 2       # Guard
        assert t["f"]["action_id"][1] == 306 and t["f"]["guard"][1] == 2 - i
        assert t["f"]["hitstun"].tolist() == [12, 12]
        assert t["reward"] == f32(0.3)
    t = blocks[3]
    assert (t["events"] >> 4) & 3 == 3           # GuardBreak
    assert t["f"]["guard"][1] == 0 and t["f"]["hitstun"].tolist() == [30, 30]
    assert t["f"]["action_id"][1] == 306 and t["f"]["reserve_id"][1] == 310
    assert t["f"]["shake"][1] == 6               # 30/3 capped at 6
    assert t["reward"] == 0                      # guard did not drop (already 0)
    i0 = next(i for i, x in enumerate(tr) if x is t)
    # reserve fires when stun reaches 0 after the decrement: 30 frames later
    assert tr[i0 + 29]["f"]["action_id"][1] == 306
    assert tr[i0 + 30]["f"]["action_id"][1] == 310 and tr[i0 + 30]["f"]["action_frame"][1] == 0
    assert tr[i0 + 30 + 35]["f"]["action_id"][1] == 310 and tr[i0 + 30 + 35]["f"]["action_frame"][1] == 35
    assert tr[i0 + 30 + 36]["f"]["action_id"][1] != 310
    # unblocked hit at guard 0: clamps at 0, normal damage (App. B-6)
    assert all(x["terminated"] == 0 for x in tr)


def test_c8_cancel_into_special_kills(oracle):
    b = make(oracle)
    place(b, -1.0, 1.2)
    # press A (N_ATTACK), press again on action frame 2 (buffer window 1-3)
    tr = run(b, [A, 0, A] + [0] * 40)
    assert tr[2]["f"]["buffer_id"][0] == 110
    assert tr[4]["events"] & 1                   # N_ATTACK connects on frame 4
    # stun 12: frames 5..16 frozen; on the frame stun hits 0 (4+12) the buffered special starts
    assert tr[4 + 11]["f"]["action_id"][0] == 100
    assert tr[4 + 12]["f"]["action_id"][0] == 110 and tr[4 + 12]["f"]["action_frame"][0] == 0
    k = next(i for i, t in enumerate(tr) if t["terminated"])
    t = tr[k]
    assert t["f"]["vital"].tolist() == [1, 0] and t["f"]["action_id"][1] == 500
    assert t["battle_over"] == 1
    assert t["obs"][3] == 0 and t["obs"][5] == 0  # DEAD -> STAND remap zeroes move_frame (footsies.py:538-552)
    # dense reward: 0.3 on first hit, special takes guard 2->1 (another 0.3) and terminal = 1 - cumulative
    assert abs(float(sum(x["reward_f64"] for x in tr[: k + 1])) - 1.0) < 1e-12
    assert t["f"]["action_frame"][0] == 11       # N_SPECIAL real hitbox starts on frame 11
    # whiff cancel is not allowed (canCancelOnWhiff 0): same inputs out of range never produce a special
    b = make(oracle)
    tr = run(b, [A, 0, A] + [0] * 40)
    assert all(t["f"]["action_id"][0] != 110 for t in tr)


def test_attack_press_during_normal_requests_special_only_inside_window(oracle):
    # a press during hitstun finds the action frozen on frame 4 -> execute window [4-5] -> buffered (Fighter.cs:492-505)
    b = make(oracle)
    place(b, -1.0, 1.2)
    tr = run(b, [A] + [0] * 9 + [A] + [0] * 20)
    assert tr[9]["f"]["buffer_id"][0] == -1 and tr[10]["f"]["buffer_id"][0] == 110
    assert tr[10]["f"]["action_frame"][0] == 4 and tr[10]["f"]["hitstun"][0] == 6
    assert tr[15]["f"]["action_id"][0] == 100 and tr[16]["f"]["action_id"][0] == 110
    # whiff: frames advance, a press on action frame 7 (outside 1-5) buffers nothing
    b = make(oracle)
    tr = run(b, [A] + [0] * 6 + [A] + [0] * 20)
    assert tr[7]["f"]["action_frame"][0] == 7
    assert all(t["f"]["buffer_id"][0] == -1 for t in tr)
    assert all(t["f"]["action_id"][0] != 110 for t in tr)


def test_c9_charge_release_special(oracle):
    # hold 59 consecutive frames then release -> N_SPECIAL; with direction held on release -> B_SPECIAL
    for held, rel, exp in ((59, 0, 110), (59, R, 115), (59, L, 115), (58, 0, None), (100, 0, 110)):
        b = make(oracle)
        tr = run(b, [A] * held + [rel] + [0] * 3)
        got = tr[held]["f"]["action_id"][0]
        if exp is None:
            assert got not in (110, 115)
        else:
            assert got == exp and tr[held]["f"]["action_frame"][0] == 0, (held, rel, got)


def test_c10_b_special_startup_invincible(oracle):
    # B_SPECIAL frames 0-5 have no hurtbox: an N_ATTACK that is active on those frames cannot hit
    b = make(oracle)
    place(b, -1.0, 1.0, p1={"action_id": 100, "action_frame": 3}, p2={"action_id": 115, "action_frame": 0})
    tr = run(b, [0] * 3)
    assert tr[0]["f"]["action_frame"].tolist() == [4, 1] and tr[0]["events"] & 3 == 0
    # next frame P2's real hitbox (frames 2-7) comes out while P1's frame-5 hitbox still finds no hurtbox
    assert tr[1]["f"]["action_frame"][1] == 2 and tr[1]["events"] & 3 == 2
    assert tr[1]["terminated"] and tr[1]["f"]["vital"].tolist() == [0, 1] and tr[1]["reward"] == f32(-1.0)
    # control: against a standing P2 the same N_ATTACK connects on its frame 4
    b = make(oracle)
    place(b, -1.0, 1.0, p1={"action_id": 100, "action_frame": 3})
    tr = run(b, [0] * 2)
    assert tr[0]["events"] & 1


def test_c11_trade_double_ko(oracle):
    # both N_SPECIAL on the same frame -> both die, Python reward +1 (App. B-9, B-10)
    b = make(oracle)
    place(b, -1.2, 1.2, p1={"action_id": 110, "action_frame": 9}, p2={"action_id": 110, "action_frame": 9})
    tr = run(b, [0] * 6)
    k = next(i for i, t in enumerate(tr) if t["terminated"])
    t = tr[k]
    assert t["f"]["vital"].tolist() == [0, 0]
    assert t["events"] & 3 == 3
    assert t["reward"] == f32(1.0)
    assert b.stats()["double_ko"] == 1


def test_proximity_guard(oracle):
    # App. B-11: holding back inside the proximity box of an attack -> GUARD_PROXIMITY next frame, no walking
    b = make(oracle)
    place(b, -1.5, 1.5)                           # distance 3.0: prox box reaches x+3.0, base hurtbox half 0.75
    tr = run(b, [A] + [0] * 10, [R] * 11)
    if hasattr(oracle, "lib"):                    # the notification itself is instrumentation of the hand-written oracle;
        assert tr[0]["events"] >> 7 & 1          # the transliterated engine shows only its effect (next line)
    assert tr[0]["f"]["is_reserve_prox"][1] == 1
    assert tr[1]["f"]["action_id"][1] == 350
    x = tr[1]["f"]["pos_x"][1]
    assert tr[2]["f"]["pos_x"][1] == x           # not walking back while in proximity
    # hitbox prox range is frames 0-5; after that P2 walks backward again
    assert tr[8]["f"]["action_id"][1] == 2


def test_stale_intro_input_leaks_into_next_episode(oracle):
    # App. B-3: the Intro frame replays the actors' last input, so input[1] at fight frame 0 is the terminal action
    b = make(oracle, autoreset=True)
    place(b, -1.2, 1.2, p1={"action_id": 110, "action_frame": 9})
    for _ in range(4):
        t = b.step([R | A], [L]).copy()[0]
        if t["terminated"]:
            break
    assert t["terminated"]
    t = b.step([0], [0]).copy()[0]               # autoreset call
    assert t["was_reset"] == 1 and t["frame"] == -1
    assert t["f"]["input0"].tolist() == [R | A, L]
    # pressing the same buttons on frame 0 is therefore NOT a fresh press: no attack, no dash
    t = b.step([R | A], [L]).copy()[0]
    assert t["f"]["action_id"].tolist() == [1, 1] and t["frame"] == 0
    # with the flag off the history is clean and the attack comes out
    b = make(oracle, autoreset=True, stale_intro_input=False)
    place(b, -1.2, 1.2, p1={"action_id": 110, "action_frame": 9})
    while not b.step([R | A], [L])[0]["terminated"]:
        pass
    b.step([0], [0])
    t = b.step([R | A], [L]).copy()[0]
    assert t["f"]["action_id"][0] == 105


def test_frame_data_matches_moves_py_table(oracle):
    # C12, via the generated python literals (the generator asserts the same against moves.py itself)
    from footsies_gym_b200 import frame_data as fd
    exp = {100: (22, 4, 2), 105: (21, 3, 3), 110: (44, 11, 4), 115: (55, 2, 6)}
    for a in fd.ACTIONS:
        if a["actionID"] in exp:
            dur, startup, active = exp[a["actionID"]]
            real = [h for h in a["hitboxes"] if not h["proximity"]]
            assert a["frameCount"] == dur
            assert min(h["se"][0] for h in real) == startup
            assert max(h["se"][1] for h in real) == startup + active - 1
    assert fd.MOVE_IDS == [0, 1, 2, 10, 11, 100, 105, 110, 115, 200, 301, 305, 306, 310, 350, 500, 510]
    assert fd.CONSTS["dashAllowFrame"] == 9 and fd.CONSTS["specialAttackHoldFrame"] == 60


def test_sparse_reward_and_loss(oracle):
    b = make(oracle, dense_reward=False)
    place(b, -1.2, 1.2, p2={"action_id": 110, "action_frame": 9})
    tr = run(b, [0] * 6)
    k = next(i for i, t in enumerate(tr) if t["terminated"])
    assert all(t["reward"] == 0 for t in tr[:k]) and tr[k]["reward"] == f32(-1.0)
    assert b.stats()["p2_wins"] == 1


def test_frame_skip_repeat_sums_reward_and_stops_at_done(oracle):
    b = make(oracle)
    place(b, -1.0, 1.2)
    t = b.step([A], [0], repeat=8).copy()[0]
    assert t["frame"] == 8                       # set_state put the battle on frame 0; 8 more frames simulated
    b2 = make(oracle)
    place(b2, -1.0, 1.2)
    tot = 0.0
    for i in range(8):
        tot += float(b2.step([A], [0])[0]["reward_f64"])
    assert float(t["reward_f64"]) == tot and tot == 0.3
    assert t["f"]["action_frame"].tolist() == b2.trace[0]["f"]["action_frame"].tolist()
