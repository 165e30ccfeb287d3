"""Expanded per-(action, frame) tables for the CUDA kernel (imported by gen_frame_data.py).

The kernel never scans ranges: every lookup the reference does through ActionData.Get*Data
(ActionData.cs:87-168) is resolved here, once, for every (action, frame) pair, into one 16-byte row
plus a few tiny geometry tables.  All fp32 constants are computed with numpy.float32 so they carry
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
DEAD_ROWS = 52  # DEAD (500 frames) is constant from frame 51 on


def fbits(v):
    return int(np.array([v], dtype=np.float32).view(np.uint32)[0])


def first_match(items, frame):
    for it in items:
        if it["se"][0] <= frame <= it["se"][1]:
            return it
    return None


def all_matches(items, frame):
    return [it for it in items if it["se"][0] <= frame <= it["se"][1]]


def build(consts, attacks, actions):
    dt = f32(consts["fixedDeltaTime"])
    base_hurt = tuple(consts["baseHurtBoxRect"])
    base_push = tuple(consts["basePushBoxRect"])
    idx_of = {a["actionID"]: i for i, a in enumerate(actions)}

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

    rows, action_info = [], []
    for a in actions:
        aid = a["actionID"]
        nrows = min(a["frameCount"], DEAD_ROWS) if aid == 500 else a["frameCount"]
        kinds = {KIND_OF_ATTACK[h["attackID"]] for h in a["hitboxes"]}
        assert len(kinds) <= 1
        kind = kinds.pop() if kinds else 0
        base = len(rows)
        for fr in range(nrows):
            dx, vel, flags = f32(0), f32(0), 0
            if aid == 1:    # FORWARD: x += forwardMoveSpeed * sign * dt, velocity_x untouched (Fighter.cs:298-302)
                dx = f32(consts["forwardMoveSpeed"]) * dt
            elif aid == 2:  # BACKWARD: x -= backwardMoveSpeed * sign * dt (Fighter.cs:303-307)
                dx = -(f32(consts["backwardMoveSpeed"]) * dt)
            else:
                m = first_match(a["movements"], fr)
                if m is not None:
                    vel = f32(m["velocity_x"])
                    dx = vel * dt if vel != 0 else f32(0)
                    flags |= 1
            hb = all_matches(a["hitboxes"], fr)
            assert len([h for h in hb if h["proximity"]]) <= 1 and len([h for h in hb if not h["proximity"]]) <= 1
            if any(h["proximity"] for h in hb):
                flags |= 2
            if any(not h["proximity"] for h in hb):
                flags |= 4
            for c in all_matches(a["cancels"], fr):
                assert c["actionID"] == [110] and (c["buffer"] or c["execute"])
                flags |= 8
            hu = all_matches(a["hurtboxes"], fr)
            assert len(hu) <= 2, "kernel keeps two hurtbox slots"
            ids = [hurt_id(base_hurt if h["useBaseRect"] else tuple(h["rect"])) for h in hu] + [0, 0]
            flags |= ids[0] << 4 | ids[1] << 8
            p = first_match(a["pushboxes"], fr)
            assert p is not None, (a["actionName"], fr)
            flags |= push_id(base_push if p["useBaseRect"] else tuple(p["rect"])) << 12
            rows.append((fbits(dx), fbits(vel), flags, 0))
        assert a["frameCount"] < 512 and base < 1024 and nrows - 1 < 64
        info = (a["frameCount"] | a["alwaysCancelable"] << 9 | (1 if a["type"] == 3 else 0) << 10 | kind << 11
                | base << 14 | (nrows - 1) << 24)
        action_info.append(info)
        assert not a["isLoop"] or aid == 510
    assert len(hurt_tab) <= 16 and len(push_tab) <= 8

    # y-overlap masks: bit j set <=> hurtbox id j overlaps this hitbox in y (BoxBase.Overlaps c3 && c4)
    hit_rows = []
    for k in (1, 2, 3, 4):
        for p in (0, 1):
            x, y, w, h = (f32(v) for v in hit_tab[(k, p)])
            ymask = 0
            for j in range(1, len(hurt_tab)):
                hx, hy, hw, hh = (f32(v) for v in hurt_tab[j])
                c3 = (hy + hh) >= y
                c4 = hy <= (y + h)
                if c3 and c4:
                    ymask |= 1 << j
            hit_rows.append((fbits(x), fbits(w / f32(2)), ymask, 0))
    hurt_rows = [(0, 0)] + [(fbits(f32(r[0])), fbits(f32(r[2]) / f32(2))) for r in hurt_tab[1:]]
    for r in hurt_tab[1:]:
        assert r[1] >= 0 and r[3] > 0
    push_rows = [(fbits(f32(r[0])), fbits(f32(r[2]))) for r in push_tab]
    for r in push_tab:  # the Rect.Overlaps y test (BattleCore.cs:488) is then always true
        assert r[1] == 0 and r[3] > 0

    atk_rows = [0]
    by_kind = {KIND_OF_ATTACK[t["attackID"]]: t for t in attacks}
    for k in (1, 2, 3, 4):
        t = by_kind[k]
        assert t["numberOfHit"] == 1 and t["guardHealthDamage"] == 1 and t["vitalHealthDamage"] in (0, 1)
        assert max(t["hitStunFrame"], t["guardStunFrame"], t["guardBreakStunFrame"]) < 32
        atk_rows.append(idx_of[t["damageActionID"]] | idx_of[t["guardActionID"]] << 5 | t["vitalHealthDamage"] << 10
                        | t["hitStunFrame"] << 11 | t["guardStunFrame"] << 16 | t["guardBreakStunFrame"] << 21)

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
    return dict(rows=rows, action_info=action_info, hit_rows=hit_rows, hurt_rows=hurt_rows, push_rows=push_rows,
                atk_rows=atk_rows, cum_vals=vals, cum_next=cum_next, term=term, step_reward=step_reward,
                hurt_tab=hurt_tab, push_tab=push_tab)


def emit(consts, attacks, actions, path):
    t = build(consts, attacks, actions)
    o = []
    w = o.append
    w("/* GENERATED by tools/gen_frame_data.py (gen_kernel_tables.py) from the reference's F00 frame data -- do not edit.")
    w(" * Expanded per-(action, frame) tables for the CUDA kernel; see tools/gen_kernel_tables.py for the layout. */")
    w("#ifndef FOOTSIES_B200_FRAME_TABLES_H")
    w("#define FOOTSIES_B200_FRAME_TABLES_H")
    w("#define FT_NUM_ACTIONS %d" % len(actions))
    w("#define FT_NUM_ROWS %d" % len(t["rows"]))
    w("#define FT_NUM_HURT %d" % len(t["hurt_rows"]))
    w("#define FT_NUM_PUSH %d" % len(t["push_rows"]))
    w("#define FT_NUM_CUM %d" % len(t["cum_vals"]))
    for i, a in enumerate(actions):
        w("#define FT_IDX_%s %d" % (a["actionName"], i))
    w("/* action idx -> CommonActionID (Fighter.cs:42-61) in moves.py order */")
    w("#define FT_ACTION_IDS_INIT {%s}" % ", ".join(str(a["actionID"]) for a in actions))
    w("/* per action: frameCount[0:9) | alwaysCancelable[9] | Type==Guard[10] | attack kind[11:14) | row base[14:24) | rows-1[24:30) */")
    w("#define FT_ACTION_INFO_INIT {%s}" % ", ".join("0x%08xu" % v for v in t["action_info"]))
    w("/* per (action, frame) row: {dx = fl(v*dt) bits, velocity_x bits, flags, 0};")
    w(" * flags: has_movement[0] | prox hitbox[1] | real hitbox[2] | cancel->110 window[3] | hurt id0[4:8) | hurt id1[8:12) | push id[12:15) */")
    w("#define FT_ROWS_INIT { \\")
    for r in t["rows"]:
        w("  {0x%08xu, 0x%08xu, 0x%08xu, 0x%08xu}, \\" % r)
    w("}")
    w("/* hit boxes [(kind-1)*2 + real]: {centre offset bits, width/2 bits, y-overlap mask over hurt ids, 0} */")
    w("#define FT_HIT_INIT {%s}" % ", ".join("{0x%08xu, 0x%08xu, 0x%08xu, 0x%08xu}" % r for r in t["hit_rows"]))
    w("/* hurt boxes [id]: {centre offset bits, width/2 bits}; id 0 = none.  %s */" % (t["hurt_tab"][1:],))
    w("#define FT_HURT_INIT {%s}" % ", ".join("{0x%08xu, 0x%08xu}" % r for r in t["hurt_rows"]))
    w("/* push boxes [id]: {centre offset bits, width bits}.  %s */" % (t["push_tab"],))
    w("#define FT_PUSH_INIT {%s}" % ", ".join("{0x%08xu, 0x%08xu}" % r for r in t["push_rows"]))
    w("/* attack [kind]: damageAction idx[0:5) | guardAction idx[5:10) | vitalDamage[10] | hitStun[11:16) | guardStun[16:21) | breakStun[21:26) */")
    w("#define FT_ATTACK_INIT {%s}" % ", ".join("0x%08xu" % v for v in t["atk_rows"]))
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
