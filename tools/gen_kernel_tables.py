"""Expanded per-(action, frame) tables for the CUDA kernel (imported by gen_frame_data.py).

The kernel never scans ranges: every lookup the reference does through ActionData.Get*Data
(ActionData.cs:87-168) is resolved here, once, for every (action, frame) pair, into one 16-byte row of a
dense [action][64 frames] table (so that the low 11 bits of the packed fighter word ARE the row index)
plus a few tiny geometry tables whose byte offsets are pre-positioned inside the row words.  All fp32 constants are computed with numpy.float32 so they carry
exactly the roundings the scalar code would produce at run time:

  row.dx   = fl(velocity_x * dt)           (Fighter.cs:300-316; the sign flip for P2 is exact)
  hit/hurt = (centre offset, width/2)      (Fighter.cs:12-15, 706-719; width/2 is exact)
  ymask    = which hurtboxes overlap a hitbox in y (y never changes: position.y == 0 always),
             using fl(y + height) on both sides (Fighter.cs:14-15, 21-22)
"""
import numpy as np

f32 = np.float32

# attackID -> attack "kind" 1..4 used by the kernel
KIND_OF_ATTACK = {1: 1, 2: 2, 10: 3, 11: 4}


def fbits(v):
    return int(np.array([v], dtype=np.float32).view(np.uint32)[0])


def first_match(items, frame):
    for it in items:
        if it["se"][0] <= frame <= it["se"][1]:
            return it
    return None


def all_matches(items, frame):
    return [it for it in items if it["se"][0] <= frame <= it["se"][1]]


def build_dash_fsm():
    """Dash detection (Fighter.CheckForwardDashInput / CheckBackwardDashInput, Fighter.cs:585-635) as a finite automaton
    over the Left/Right bits of each frame's input.

    The checks look, among the 8 most recent past frames, for the most recent one with a direction held; it must hold
    only the direction being pressed now, and one of the 8 frames before it must be neutral.  All they can ever read
    of the past is therefore: how many frames ago that frame was (`since`, 0 = the previous frame), which directions
    it held (`lastdir`, Left 1 | Right 2) and how long the unbroken run of direction-held frames ending there is
    (`runlen`: the frame before the run is neutral, so the neutral-frame condition is runlen <= 8).  Once `since`
    exceeds 7 nothing can be read any more: one COLD state (id 0, also the state of a cleared history).

    Returns (states, table): states[id] = (since, runlen, lastdir) with COLD = (8, 0, 0); table[id][d] =
    next id | dash-by-Left << 8 | dash-by-Right << 9 for the new frame's direction bits d."""
    states = [(8, 0, 0)]
    for since in range(8):
        for runlen in range(1, 10):       # 9 = "9 or more"
            for lastdir in (1, 2, 3):
                states.append((since, runlen, lastdir))
    index = {st: i for i, st in enumerate(states)}
    table = []
    for (since, runlen, lastdir) in states:
        row = []
        cold = since >= 8
        for d in range(4):
            dash = 0
            for bit in (1, 2):            # pressing Left (1) / Right (2) this frame
                pressed = (d & bit) and not (not cold and since == 0 and (lastdir & bit))
                if pressed and not cold and since <= 7 and lastdir == bit and runlen <= 8:
                    dash |= bit
            if d:
                nxt = (0, min(runlen + 1, 9) if (not cold and since == 0) else 1, d)
            elif cold or since + 1 >= 8:
                nxt = (8, 0, 0)
            else:
                nxt = (since + 1, runlen, lastdir)
            row.append(index[nxt] | dash << 8)
        table.append(row)
    assert len(states) <= 256
    return states, table


def build_attack_run_lut():
    """Attack-button run length (saturating at 59 = specialAttackHoldFrame - 1) as a lookup: [run * 8 + input] ->
    new run | special << 6 | attack-down << 7.  special = Attack released after being held on input[1..59]
    (Fighter.CheckSpecialAttackInput, Fighter.cs:569-583); attack-down = IsAttackInput(inputDown[0]) (:184-185)."""
    lut = []
    for run in range(64):
        for inp in range(8):
            a = (inp >> 2) & 1
            r = min(run, 59)
            new = min(r + 1, 59) if a else 0
            special = int((not a) and r >= 59)
            down = int(a and r == 0)
            lut.append(new | special << 6 | down << 7)
    return lut


# request LUT index bits (frame_logic.cuh assembles them with three shifts): Left/Right now [0:2) |
# isReserveProximityGuard [2] | special [3] | attack-down [4] | dash-by-Left [5] | dash-by-Right [6] | carry END [7] |
# carry ALWAYS [8] | carry NORMAL [9]
REQ_FREE, REQ_WANT_BUFFER, REQ_ENDED = 16, 32, 64


def build_request_lut(idx_of):
    """Fighter.UpdateActionRequest (Fighter.cs:201-286) with the RequestAction chain (Fighter.cs:472-510) collapsed, as a
    lookup per side: [side][index] = requested action index | FREE (the action ended or is alwaysCancelable, so the
    first request of the chain wins) | WANT_BUFFER (not free and the request is N_SPECIAL: buffered if the frame is
    inside a cancel window) | ENDED.  Priority: special release (-> B_SPECIAL if a direction is held, else N_SPECIAL),
    attack press (-> N_SPECIAL while in N/B_ATTACK, else B_ATTACK / N_ATTACK), forward dash, backward dash, movement
    (both or no direction -> STAND, forward -> FORWARD, back -> GUARD_PROXIMITY if reserved else BACKWARD)."""
    A = idx_of
    luts = []
    for side in (0, 1):                       # P1 faces right: forward = Right (2); P2 faces left: forward = Left (1)
        fwd_bit, back_bit = (2, 1) if side == 0 else (1, 2)
        lut = []
        for idx in range(1024):
            lr = idx & 3
            rprox, special, down = (idx >> 2) & 1, (idx >> 3) & 1, (idx >> 4) & 1
            dash_l, dash_r = (idx >> 5) & 1, (idx >> 6) & 1
            ended, always, normal = (idx >> 7) & 1, (idx >> 8) & 1, (idx >> 9) & 1
            dash_f, dash_b = (dash_r, dash_l) if side == 0 else (dash_l, dash_r)
            dirn = 1 if lr else 0
            in_normal = normal and not ended
            fwd, back = bool(lr & fwd_bit), bool(lr & back_bit)
            if special:
                req = A["N_SPECIAL"] + dirn               # B_SPECIAL = N_SPECIAL + 1
            elif down:
                req = A["N_SPECIAL"] if in_normal else A["N_ATTACK"] + dirn
            elif dash_f:
                req = A["DASH_FORWARD"]
            elif dash_b:
                req = A["DASH_BACKWARD"]
            elif fwd and back:
                req = A["STAND"]
            elif fwd:
                req = A["FORWARD"]
            elif back:
                req = A["GUARD_PROXIMITY"] if rprox else A["BACKWARD"]
            else:
                req = A["STAND"]
            free = ended or always
            want_buffer = (not free) and req == A["N_SPECIAL"]
            lut.append(req | (REQ_FREE if free else 0) | (REQ_WANT_BUFFER if want_buffer else 0) | (REQ_ENDED if ended else 0))
        luts.append(lut)
    assert A["B_SPECIAL"] == A["N_SPECIAL"] + 1 and A["B_ATTACK"] == A["N_ATTACK"] + 1
    return luts


ROW_FRAMES = 64        # rows per action; the 6-bit frame field of the packed fighter word indexes them directly

# row.z bit layout
Z_HAS_MOVEMENT, Z_PROX, Z_REAL, Z_CANCEL, Z_GUARDING = 1, 2, 4, 8, 16
Z_BOXCFG_SHIFT = 5     # 4 bits: box configuration id, positioned so that (z & 0x1e0) == id * 32 (byte offset)
Z_YMASK0_SHIFT = 16    # 8 bits: which hit boxes [(kind-1)*2 + real] overlap hurt box 0 in y
Z_YMASK1_SHIFT = 24    # 8 bits: same for hurt box 1
# row.w bit layout
W_KIND_SHIFT = 5       # 3 bits: attack kind, positioned so that (w & 0xe0) == kind * 32 (byte offset)
W_CARRY_END, W_CARRY_ALWAYS, W_CARRY_NORMAL = 1 << 28, 1 << 29, 1 << 30   # same bits as the packed fighter word


def build(consts, attacks, actions):
    dt = f32(consts["fixedDeltaTime"])
    base_hurt = tuple(consts["baseHurtBoxRect"])
    base_push = tuple(consts["basePushBoxRect"])
    idx_of = {a["actionID"]: i for i, a in enumerate(actions)}
    name_of = {a["actionID"]: a["actionName"] for a in actions}

    hurt_tab = [None]          # id 0 = none
    push_tab = []

    def hurt_id(r):
        if r not in hurt_tab:
            hurt_tab.append(r)
        return hurt_tab.index(r)

    def push_id(r):
        if r not in push_tab:
            push_tab.append(r)
        return push_tab.index(r)

    hurt_id(base_hurt)
    push_id(base_push)

    # hit boxes: [kind][0 = proximity, 1 = real]
    hit_tab = {}
    for a in actions:
        for h in a["hitboxes"]:
            k = KIND_OF_ATTACK[h["attackID"]]
            key = (k, 0 if h["proximity"] else 1)
            assert key not in hit_tab or hit_tab[key] == h["rect"], "kernel assumes one prox + one real box per attack"
            hit_tab[key] = h["rect"]
    assert sorted(hit_tab) == [(k, p) for k in (1, 2, 3, 4) for p in (0, 1)]

    def ymask_of(hid):
        """bit (kind-1)*2 + real set <=> hurt box `hid` overlaps that hit box in y (BoxBase.Overlaps c3 && c4)"""
        if hid == 0:
            return 0
        hx, hy, hw, hh = (f32(v) for v in hurt_tab[hid])
        m = 0
        for k in (1, 2, 3, 4):
            for p in (0, 1):
                x, y, w, h = (f32(v) for v in hit_tab[(k, p)])
                if (hy + hh) >= y and hy <= (y + h):
                    m |= 1 << ((k - 1) * 2 + p)
        return m

    # pass 1: per (action, frame) raw facts
    raw = {}
    cfg_tab = []
    action_info = []
    for ai, a in enumerate(actions):
        aid = a["actionID"]
        kinds = {KIND_OF_ATTACK[h["attackID"]] for h in a["hitboxes"]}
        assert len(kinds) <= 1
        kind = kinds.pop() if kinds else 0
        assert a["frameCount"] <= ROW_FRAMES - 1 or aid == 500, a["actionName"]
        assert not a["isLoop"] or aid == 510
        for fr in range(ROW_FRAMES):
            src = min(fr, a["frameCount"] - 1)          # frames past the end repeat the last frame (never read)
            dx, vel, z = f32(0), f32(0), 0
            if aid == 1:    # FORWARD: x += forwardMoveSpeed * sign * dt, velocity_x untouched (Fighter.cs:298-302)
                dx = f32(consts["forwardMoveSpeed"]) * dt
            elif aid == 2:  # BACKWARD: x -= backwardMoveSpeed * sign * dt (Fighter.cs:303-307)
                dx = -(f32(consts["backwardMoveSpeed"]) * dt)
            else:
                m = first_match(a["movements"], src)
                if m is not None:
                    vel = f32(m["velocity_x"])
                    dx = vel * dt if vel != 0 else f32(0)
                    z |= Z_HAS_MOVEMENT
            hb = all_matches(a["hitboxes"], src)
            assert len([h for h in hb if h["proximity"]]) <= 1 and len([h for h in hb if not h["proximity"]]) <= 1
            if any(h["proximity"] for h in hb):
                z |= Z_PROX
            if any(not h["proximity"] for h in hb):
                z |= Z_REAL
            for c in all_matches(a["cancels"], src):
                assert c["actionID"] == [110] and (c["buffer"] or c["execute"])
                z |= Z_CANCEL
            # Fighter.NotifyDamaged blocks when the victim is in BACKWARD or any Type == Guard action (Fighter.cs:366-369)
            if aid == 2 or a["type"] == 3:
                z |= Z_GUARDING
            hu = all_matches(a["hurtboxes"], src)
            assert len(hu) <= 2, "kernel keeps two hurtbox slots"
            ids = [hurt_id(base_hurt if h["useBaseRect"] else tuple(h["rect"])) for h in hu] + [0, 0]
            pb = first_match(a["pushboxes"], src)
            assert pb is not None, (a["actionName"], fr)
            cfg = (ids[0], ids[1], push_id(base_push if pb["useBaseRect"] else tuple(pb["rect"])))
            if cfg not in cfg_tab:
                cfg_tab.append(cfg)
            raw[(ai, fr)] = (dx, vel, z, cfg, kind)
        action_info.append(a["frameCount"] | a["alwaysCancelable"] << 9 | (1 if a["type"] == 3 else 0) << 10 | kind << 11)
    assert len(cfg_tab) <= 16 and len(hurt_tab) <= 16 and len(push_tab) <= 8
    for r in hurt_tab[1:]:
        assert r[1] >= 0 and r[3] > 0
    for r in push_tab:  # the Rect.Overlaps y test (BattleCore.cs:488) is then always true
        assert r[1] == 0 and r[3] > 0

    # pass 2: rows
    rows = []
    for ai, a in enumerate(actions):
        is_normal = name_of[a["actionID"]] in ("N_ATTACK", "B_ATTACK")
        for fr in range(ROW_FRAMES):
            dx, vel, z, cfg, kind = raw[(ai, fr)]
            z |= cfg_tab.index(cfg) << Z_BOXCFG_SHIFT
            z |= ymask_of(cfg[0]) << Z_YMASK0_SHIFT | ymask_of(cfg[1]) << Z_YMASK1_SHIFT
            w = kind << W_KIND_SHIFT
            # carry bits describe what the NEXT frame's request logic needs to know about this action:
            #   END:    the action is over once the frame counter increments (frame + 1 >= frameCount, Fighter.cs:90)
            #   ALWAYS: alwaysCancelable;   NORMAL: N_ATTACK / B_ATTACK (attack press cancels into N_SPECIAL, Fighter.cs:246-252)
            if fr + 1 >= a["frameCount"]:
                w |= W_CARRY_END
            if a["alwaysCancelable"]:
                w |= W_CARRY_ALWAYS
            if is_normal:
                w |= W_CARRY_NORMAL
            rows.append((fbits(dx), fbits(vel), z, w))

    # box configurations: {hurt0 centre, hurt0 width/2, hurt1 centre, hurt1 width/2, push centre, push width, 0, 0}
    cfg_rows = []
    for (h0, h1, pb) in cfg_tab:
        def hb(hid):
            if hid == 0:
                return (0, 0)
            r = hurt_tab[hid]
            return (fbits(f32(r[0])), fbits(f32(r[2]) / f32(2)))
        pr = push_tab[pb]
        cfg_rows.append(hb(h0) + hb(h1) + (fbits(f32(pr[0])), fbits(f32(pr[2])), 0, 0))

    # attacks [kind]: {prox centre, prox width/2, real centre, real width/2, prox y-bit pair, real y-bit pair, result word, 0}
    by_kind = {KIND_OF_ATTACK[t["attackID"]]: t for t in attacks}
    atk_rows = [(0,) * 8]
    for k in (1, 2, 3, 4):
        t = by_kind[k]
        assert t["numberOfHit"] == 1 and t["guardHealthDamage"] == 1 and t["vitalHealthDamage"] in (0, 1)
        assert max(t["hitStunFrame"], t["guardStunFrame"], t["guardBreakStunFrame"]) < 32
        result = (idx_of[t["damageActionID"]] | idx_of[t["guardActionID"]] << 5 | t["vitalHealthDamage"] << 10
                  | t["hitStunFrame"] << 11 | t["guardStunFrame"] << 16 | t["guardBreakStunFrame"] << 21)
        px, _, pw, _ = (f32(v) for v in hit_tab[(k, 0)])
        rx, _, rw, _ = (f32(v) for v in hit_tab[(k, 1)])
        pbit, rbit = 1 << ((k - 1) * 2), 1 << ((k - 1) * 2 + 1)
        atk_rows.append((fbits(px), fbits(pw / f32(2)), fbits(rx), fbits(rw / f32(2)),
                         pbit << Z_YMASK0_SHIFT | pbit << Z_YMASK1_SHIFT, rbit << Z_YMASK0_SHIFT | rbit << Z_YMASK1_SHIFT,
                         result, 0))

    # dense-reward automaton (footsies.py:388-405): Python accumulates 0.3 steps in float64; the set of
    # reachable cumulative values is tiny, so the kernel carries an index and the doubles live in a table.
    vals = [0.0]
    trans = {}
    seen = {(0.0, 0, 0)}
    todo = [(0.0, 0, 0)]
    while todo:
        c, m, p = todo.pop()
        for dm, dp in ((1, 0), (0, 1), (1, 1)):
            if m + dm > 3 or p + dp > 3:
                continue
            r = 0.0
            if dm:
                r -= 0.3
            if dp:
                r += 0.3
            n = c + r
            if n not in vals:
                vals.append(n)
            trans[(vals.index(c), dm, dp)] = vals.index(n)
            if (n, m + dm, p + dp) not in seen:
                seen.add((n, m + dm, p + dp))
                todo.append((n, m + dm, p + dp))
    assert len(vals) <= 16
    cum_next = []   # [idx][code]  code: 0 none, 1 P1 guard dropped, 2 P2 guard dropped, 3 both
    term = []       # [idx_after][code][p2_dead]  = step_reward + ((+1 | -1) - cum_after)   (float64)
    def step_r(code):
        r = 0.0
        if code & 1:
            r -= 0.3
        if code & 2:
            r += 0.3
        return r

    for i, c in enumerate(vals):
        # transitions that would need more than 3 guard drops on one side are unreachable -> 0
        cum_next.append([vals.index(c + step_r(code)) if (c + step_r(code)) in vals else 0 for code in range(4)])
        for (j, dm, dp), k in trans.items():
            if j == i:
                assert cum_next[i][dm | dp << 1] == k
        rowt = []
        for code in range(4):
            r = 0.0
            if code & 1:
                r -= 0.3
            if code & 2:
                r += 0.3
            rowt.append([r + ((-1) - c), r + (1 - c)])
        term.append(rowt)
    step_reward = []
    for code in range(4):
        r = 0.0
        if code & 1:
            r -= 0.3
        if code & 2:
            r += 0.3
        step_reward.append(r)
    dash_states, dash_table = build_dash_fsm()
    req_lut = build_request_lut({a["actionName"]: i for i, a in enumerate(actions)})
    return dict(rows=rows, action_info=action_info, cfg_rows=cfg_rows, cfg_tab=cfg_tab, atk_rows=atk_rows, cum_vals=vals,
                cum_next=cum_next, term=term, step_reward=step_reward, hurt_tab=hurt_tab, push_tab=push_tab,
                dash_states=dash_states, dash_table=dash_table, arun_lut=build_attack_run_lut(), req_lut=req_lut)



def emit(consts, attacks, actions, path):
    t = build(consts, attacks, actions)
    o = []
    w = o.append
    w("/* GENERATED by tools/gen_frame_data.py (gen_kernel_tables.py) from the reference's F00 frame data -- do not edit.")
    w(" * Expanded per-(action, frame) tables for the CUDA kernel; see tools/gen_kernel_tables.py for the layout. */")
    w("#ifndef FOOTSIES_B200_FRAME_TABLES_H")
    w("#define FOOTSIES_B200_FRAME_TABLES_H")
    w("#define FT_NUM_ACTIONS %d" % len(actions))
    w("#define FT_ROW_FRAMES %d" % ROW_FRAMES)
    w("#define FT_NUM_ROWS %d" % len(t["rows"]))
    w("#define FT_NUM_BOXCFG %d" % len(t["cfg_rows"]))
    w("#define FT_NUM_CUM %d" % len(t["cum_vals"]))
    for i, a in enumerate(actions):
        w("#define FT_IDX_%s %d" % (a["actionName"], i))
    w("/* row.z / row.w bit layout (tools/gen_kernel_tables.py) */")
    for name in ("Z_HAS_MOVEMENT", "Z_PROX", "Z_REAL", "Z_CANCEL", "Z_GUARDING", "Z_BOXCFG_SHIFT", "Z_YMASK0_SHIFT",
                 "Z_YMASK1_SHIFT", "W_KIND_SHIFT", "W_CARRY_END", "W_CARRY_ALWAYS", "W_CARRY_NORMAL"):
        w("#define FT_%s 0x%xu" % (name, globals()[name]))
    w("/* action idx -> CommonActionID (Fighter.cs:42-61) in moves.py order */")
    w("#define FT_ACTION_IDS_INIT {%s}" % ", ".join(str(a["actionID"]) for a in actions))
    w("/* per action (host side only): frameCount[0:9) | alwaysCancelable[9] | Type==Guard[10] | attack kind[11:14) */")
    w("#define FT_ACTION_INFO_INIT {%s}" % ", ".join("0x%08xu" % v for v in t["action_info"]))
    w("/* rows[action * 64 + frame] = {dx = fl(v*dt) bits, velocity_x bits, z, w};")
    w(" * z: has_movement[0] | prox hitbox[1] | real hitbox[2] | cancel->110 window[3] | blocks when hit[4] | box config id[5:9) |")
    w(" *    y-overlap bits of hurt box 0 [16:24) and hurt box 1 [24:32) over hit boxes (kind-1)*2+real;")
    w(" * w: attack kind[5:8) | carry END[28] ALWAYS[29] NORMAL[30] (copied into the packed fighter word) */")
    w("#define FT_ROWS_INIT { \\")
    for r in t["rows"]:
        w("  {0x%08xu, 0x%08xu, 0x%08xu, 0x%08xu}, \\" % r)
    w("}")
    w("/* box configurations [id]: {hurt0 centre, hurt0 width/2, hurt1 centre, hurt1 width/2, push centre, push width, 0, 0} (fp32 bits)")
    w(" * (hurt0 id, hurt1 id, push id) = %s" % (t["cfg_tab"],))
    w(" * hurt boxes %s" % (t["hurt_tab"][1:],))
    w(" * push boxes %s */" % (t["push_tab"],))
    w("#define FT_BOXCFG_INIT {%s}" % ", ".join("{%s}" % ", ".join("0x%08xu" % v for v in r) for r in t["cfg_rows"]))
    w("/* attacks [kind]: {prox centre, prox width/2, real centre, real width/2, prox y-bits, real y-bits, result, 0};")
    w(" * result: damageAction idx[0:5) | guardAction idx[5:10) | vitalDamage[10] | hitStun[11:16) | guardStun[16:21) | breakStun[21:26) */")
    w("#define FT_ATTACK_INIT {%s}" % ", ".join("{%s}" % ", ".join("0x%08xu" % v for v in r) for r in t["atk_rows"]))
    w("/* dash detection automaton (tools/gen_kernel_tables.py build_dash_fsm): [state][direction bits] = next state |")
    w(" * dash-by-Left << 8 | dash-by-Right << 9; state 0 = COLD (nothing readable in the last 8 frames) */")
    w("#define FT_NUM_DASH_STATES %d" % len(t["dash_states"]))
    w("#define FT_DASH_FSM_INIT {%s}" % ", ".join("{%s}" % ", ".join("0x%03x" % v for v in r) for r in t["dash_table"]))
    w("/* [state] = since | runlen << 4 | lastdir << 8 (host side: expansion to / from a Left/Right bit history) */")
    w("#define FT_DASH_STATE_INFO_INIT {%s}" % ", ".join("0x%03x" % (st[0] | st[1] << 4 | st[2] << 8) for st in t["dash_states"]))
    w("/* request lookup [side][Left/Right | rprox << 2 | special << 3 | attack-down << 4 | dash-by-Left << 5 | dash-by-Right << 6 |")
    w(" * END << 7 | ALWAYS << 8 | NORMAL << 9] = requested action | FREE 16 | WANT_BUFFER 32 | ENDED 64 (build_request_lut) */")
    w("#define FT_REQ_FREE %du" % REQ_FREE)
    w("#define FT_REQ_WANT_BUFFER %du" % REQ_WANT_BUFFER)
    w("#define FT_REQ_ENDED %du" % REQ_ENDED)
    w("#define FT_REQ_LUT_INIT {%s}" % ", ".join("{%s}" % ", ".join(str(v) for v in l) for l in t["req_lut"]))
    w("/* attack run length [run * 8 + input] = new run | special << 6 | attack-down << 7 */")
    w("#define FT_ARUN_LUT_INIT {%s}" % ", ".join(str(v) for v in t["arun_lut"]))
    w("/* dense reward automaton (footsies.py:388-405): cumulative float64 values, next index per guard-drop code, terminal reward */")
    w("#define FT_CUM_VALUES_INIT {%s}" % ", ".join(float(v).hex() for v in t["cum_vals"]))
    w("#define FT_CUM_NEXT_INIT {%s}" % ", ".join("{%s}" % ", ".join(map(str, r)) for r in t["cum_next"]))
    w("#define FT_STEP_REWARD_INIT {%s}" % ", ".join(float(v).hex() for v in t["step_reward"]))
    w("/* [cum idx after][code][p2 dead] */")
    w("#define FT_TERM_REWARD_INIT {%s}" % ", ".join(
        "{%s}" % ", ".join("{%s, %s}" % (float(c[0]).hex(), float(c[1]).hex()) for c in r) for r in t["term"]))
    w("#endif")
    with open(path, "w") as f:
        f.write("\n".join(o) + "\n")


if __name__ == "__main__":
    # regenerate from the committed Python literals (no reference tree needed)
    import os
    import sys
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, repo)
    from footsies_gym_b200 import frame_data as fd
    emit(fd.CONSTS, fd.ATTACKS, fd.ACTIONS, os.path.join(repo, "footsies_gym_b200", "csrc", "frame_tables.h"))
    print("wrote footsies_gym_b200/csrc/frame_tables.h")
