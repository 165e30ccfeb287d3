#!/usr/bin/env python3
"""Build-time extractor: FOOTSIES frame data (Unity ScriptableObject YAML) -> generated tables.

Runs ONLY in the authoring container (needs /root/reference); its outputs are
committed so nothing on the GPU box ever reads the reference tree:

  oracle/frame_data.h                     range-form tables, scanned by the CPU oracle exactly the
                                          way ActionData.Get*Data scans them (ActionData.cs:87-168)
  footsies_gym_b200/csrc/frame_tables.h   per-(action,frame) expanded tables for the CUDA kernel
  footsies_gym_b200/frame_data.py         the same data as Python literals (front-end + tests)

Sources (file:line relative to /root/reference):
  Assets/Fighter/F00/F00.asset:14-31                     fighter constants
  Assets/Fighter/F00/F00_AttackDataContainer.asset:14-54 attack results
  Assets/Fighter/F00/Actions/*.asset                     17 actions
  Assets/Scenes/BattleScene.unity:273                    _battleAreaWidth
  ProjectSettings/TimeManager.asset:6                    Fixed Timestep
Cross-checked against footsies-gym/footsies_gym/moves.py:13-29 (duration / startup / active).
"""
import argparse
import importlib.util
import os
import pprint
import struct
import sys

import numpy as np
import yaml

REF = "/root/reference"
F00 = os.path.join(REF, "Assets/Fighter/F00")
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_unity_yaml(path):
    with open(path, encoding="utf-8-sig") as f:
        lines = f.read().splitlines()
    # drop '%YAML', '%TAG' and the '--- !u!NNN &id' document header
    body = [ln for ln in lines if not (ln.startswith("%") or ln.startswith("---"))]
    return yaml.safe_load("\n".join(body))


def parse_int_list_blob(blob):
    """Unity serialises List<int> as a hex string of little-endian int32s ('6e000000' == [110])."""
    if blob is None or blob == "":
        return []
    s = str(blob)
    if len(s) % 8:
        s = s.zfill((len(s) + 7) // 8 * 8)
    return [struct.unpack("<i", bytes.fromhex(s[i:i + 8]))[0] for i in range(0, len(s), 8)]


def f32(v):
    return float(np.float32(v))


def rect(d):
    return (f32(d["x"]), f32(d["y"]), f32(d["width"]), f32(d["height"]))


def se(d):
    return (int(d["startEndFrame"]["x"]), int(d["startEndFrame"]["y"]))


def load_all():
    fighter = load_unity_yaml(os.path.join(F00, "F00.asset"))["MonoBehaviour"]
    consts = dict(
        startGuardHealth=int(fighter["startGuardHealth"]),
        forwardMoveSpeed=f32(fighter["forwardMoveSpeed"]),
        backwardMoveSpeed=f32(fighter["backwardMoveSpeed"]),
        dashAllowFrame=int(fighter["dashAllowFrame"]),
        specialAttackHoldFrame=int(fighter["specialAttackHoldFrame"]),
        canCancelOnWhiff=int(fighter["canCancelOnWhiff"]),
        baseHurtBoxRect=rect(fighter["baseHurtBoxRect"]),
        basePushBoxRect=rect(fighter["basePushBoxRect"]),
    )
    scene = open(os.path.join(REF, "Assets/Scenes/BattleScene.unity")).read()
    for ln in scene.splitlines():
        if ln.strip().startswith("_battleAreaWidth:"):
            consts["battleAreaWidth"] = f32(ln.split(":")[1])
    tm = open(os.path.join(REF, "ProjectSettings/TimeManager.asset")).read()
    for ln in tm.splitlines():
        if ln.strip().startswith("Fixed Timestep:"):
            consts["fixedDeltaTime"] = f32(ln.split(":")[1])

    attacks = []
    for a in load_unity_yaml(os.path.join(F00, "F00_AttackDataContainer.asset"))["MonoBehaviour"]["attackDataList"]:
        attacks.append({k: (a[k] if k == "attackName" else int(a[k])) for k in (
            "attackID", "attackName", "damageActionID", "guardActionID", "numberOfHit",
            "vitalHealthDamage", "guardHealthDamage", "hitStunFrame", "guardStunFrame",
            "guardBreakStunFrame")})
    attacks.sort(key=lambda a: a["attackID"])

    actions = []
    adir = os.path.join(F00, "Actions")
    for fn in sorted(os.listdir(adir)):
        if not fn.endswith(".asset"):
            continue
        m = load_unity_yaml(os.path.join(adir, fn))["MonoBehaviour"]
        actions.append(dict(
            actionID=int(m["actionID"]), actionName=str(m["actionName"]), type=int(m["Type"]),
            frameCount=int(m["frameCount"]), isLoop=int(m.get("isLoop", 0)), loopFromFrame=int(m.get("loopFromFrame", 0)),  # absent => C# default
            alwaysCancelable=int(m.get("alwaysCancelable", 0)),
            hitboxes=[dict(se=se(h), rect=rect(h["rect"]), attackID=int(h["attackID"]),
                           proximity=int(h["proximity"])) for h in (m["hitboxes"] or [])],
            hurtboxes=[dict(se=se(h), rect=rect(h["rect"]), useBaseRect=int(h["useBaseRect"]))
                       for h in (m["hurtboxes"] or [])],
            pushboxes=[dict(se=se(h), rect=rect(h["rect"]), useBaseRect=int(h["useBaseRect"]))
                       for h in (m["pushboxes"] or [])],
            movements=[dict(se=se(h), velocity_x=f32(h["velocity_x"])) for h in (m["movements"] or [])],
            cancels=[dict(se=se(h), buffer=int(h["buffer"]), execute=int(h["execute"]),
                          actionID=parse_int_list_blob(h["actionID"])) for h in (m["cancels"] or [])],
        ))
    actions.sort(key=lambda a: a["actionID"])
    return consts, attacks, actions


def crosscheck_moves_py(actions):
    """moves.py is the only per-move known-answer table the reference ships."""
    spec = importlib.util.spec_from_file_location(
        "ref_moves", os.path.join(REF, "footsies-gym/footsies_gym/moves.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ref_moves"] = mod
    spec.loader.exec_module(mod)
    order = [m.value.id for m in mod.FOOTSIES_MOVE_INDEX_TO_MOVE]
    assert order == [a["actionID"] for a in actions], "action order differs from moves.py"
    by_id = {a["actionID"]: a for a in actions}
    for mv in mod.FootsiesMove:
        a = by_id[mv.value.id]
        assert a["frameCount"] == mv.value.duration, (mv, a["frameCount"])
        real = [h for h in a["hitboxes"] if not h["proximity"]]
        if mv.value.active:
            first = min(h["se"][0] for h in real)
            last = max(h["se"][1] for h in real)
            assert first == mv.value.startup and last == mv.value.startup + mv.value.active - 1, mv
            assert mv.value.startup + mv.value.active + mv.value.recovery == mv.value.duration, mv
        else:
            assert not real, mv
    return [m.name for m in mod.FOOTSIES_MOVE_INDEX_TO_MOVE]


def cf(v):
    """float -> C float literal that round-trips the fp32 value."""
    s = "%.9g" % np.float32(v)
    if "." not in s and "e" not in s and "inf" not in s:
        s += ".0"
    return s + "f"


def emit_oracle_header(consts, attacks, actions, path):
    o = []
    w = o.append
    w("/* GENERATED by tools/gen_frame_data.py from /root/reference/Assets/Fighter/F00 -- do not edit.")
    w(" * Range-form frame data, scanned by the oracle the way ActionData.Get*Data does (ActionData.cs:87-168). */")
    w("#ifndef FOOTSIES_ORACLE_FRAME_DATA_H")
    w("#define FOOTSIES_ORACLE_FRAME_DATA_H")
    w("typedef struct { float x, y, width, height; } fd_rect;")
    w("typedef struct { int start, end; fd_rect rect; int attackID; int proximity; } fd_hitbox;")
    w("typedef struct { int start, end; fd_rect rect; int useBaseRect; } fd_box;")
    w("typedef struct { int start, end; float velocity_x; } fd_movement;")
    w("typedef struct { int start, end; int buffer, execute; int n_ids; int ids[4]; } fd_cancel;")
    w("typedef struct { int actionID; const char *name; int type; int frameCount; int isLoop; int loopFromFrame;")
    w("  int alwaysCancelable;")
    w("  int n_hitboxes; const fd_hitbox *hitboxes; int n_hurtboxes; const fd_box *hurtboxes;")
    w("  int n_pushboxes; const fd_box *pushboxes; int n_movements; const fd_movement *movements;")
    w("  int n_cancels; const fd_cancel *cancels; } fd_action;")
    w("typedef struct { int attackID; const char *name; int damageActionID, guardActionID, numberOfHit,")
    w("  vitalHealthDamage, guardHealthDamage, hitStunFrame, guardStunFrame, guardBreakStunFrame; } fd_attack;")
    w("")
    w("/* F00.asset:14-31, BattleScene.unity:273, TimeManager.asset:6 */")
    w("#define FD_START_GUARD_HEALTH %d" % consts["startGuardHealth"])
    w("#define FD_FORWARD_MOVE_SPEED %s" % cf(consts["forwardMoveSpeed"]))
    w("#define FD_BACKWARD_MOVE_SPEED %s" % cf(consts["backwardMoveSpeed"]))
    w("#define FD_DASH_ALLOW_FRAME %d" % consts["dashAllowFrame"])
    w("#define FD_SPECIAL_ATTACK_HOLD_FRAME %d" % consts["specialAttackHoldFrame"])
    w("#define FD_CAN_CANCEL_ON_WHIFF %d" % consts["canCancelOnWhiff"])
    w("#define FD_BATTLE_AREA_WIDTH %s" % cf(consts["battleAreaWidth"]))
    w("#define FD_FIXED_DELTA_TIME %s" % cf(consts["fixedDeltaTime"]))
    w("static const fd_rect FD_BASE_HURTBOX = {%s};" % ", ".join(cf(v) for v in consts["baseHurtBoxRect"]))
    w("static const fd_rect FD_BASE_PUSHBOX = {%s};" % ", ".join(cf(v) for v in consts["basePushBoxRect"]))
    w("")
    for a in actions:
        n = a["actionName"]
        if a["hitboxes"]:
            w("static const fd_hitbox FD_%s_HIT[] = {" % n)
            for h in a["hitboxes"]:
                w("  {%d, %d, {%s}, %d, %d}," % (h["se"][0], h["se"][1], ", ".join(cf(v) for v in h["rect"]),
                                                h["attackID"], h["proximity"]))
            w("};")
        for key, tag in (("hurtboxes", "HURT"), ("pushboxes", "PUSH")):
            if a[key]:
                w("static const fd_box FD_%s_%s[] = {" % (n, tag))
                for h in a[key]:
                    w("  {%d, %d, {%s}, %d}," % (h["se"][0], h["se"][1], ", ".join(cf(v) for v in h["rect"]),
                                                 h["useBaseRect"]))
                w("};")
        if a["movements"]:
            w("static const fd_movement FD_%s_MOVE[] = {" % n)
            for h in a["movements"]:
                w("  {%d, %d, %s}," % (h["se"][0], h["se"][1], cf(h["velocity_x"])))
            w("};")
        if a["cancels"]:
            w("static const fd_cancel FD_%s_CANCEL[] = {" % n)
            for h in a["cancels"]:
                ids = h["actionID"] + [0] * (4 - len(h["actionID"]))
                w("  {%d, %d, %d, %d, %d, {%s}}," % (h["se"][0], h["se"][1], h["buffer"], h["execute"],
                                                     len(h["actionID"]), ", ".join(map(str, ids))))
            w("};")
    w("")
    w("#define FD_NUM_ACTIONS %d" % len(actions))
    w("static const fd_action FD_ACTIONS[FD_NUM_ACTIONS] = {")
    for a in actions:
        n = a["actionName"]

        def ref(key, tag):
            return ("%d, FD_%s_%s" % (len(a[key]), n, tag)) if a[key] else "0, 0"
        w('  {%d, "%s", %d, %d, %d, %d, %d, %s, %s, %s, %s, %s},' % (
            a["actionID"], n, a["type"], a["frameCount"], a["isLoop"], a["loopFromFrame"], a["alwaysCancelable"],
            ref("hitboxes", "HIT"), ref("hurtboxes", "HURT"), ref("pushboxes", "PUSH"),
            ref("movements", "MOVE"), ref("cancels", "CANCEL")))
    w("};")
    w("#define FD_NUM_ATTACKS %d" % len(attacks))
    w("static const fd_attack FD_ATTACKS[FD_NUM_ATTACKS] = {")
    for t in attacks:
        w('  {%d, "%s", %d, %d, %d, %d, %d, %d, %d, %d},' % (
            t["attackID"], t["attackName"], t["damageActionID"], t["guardActionID"], t["numberOfHit"],
            t["vitalHealthDamage"], t["guardHealthDamage"], t["hitStunFrame"], t["guardStunFrame"],
            t["guardBreakStunFrame"]))
    w("};")
    w("#endif")
    with open(path, "w") as f:
        f.write("\n".join(o) + "\n")


def emit_python(consts, attacks, actions, names, path):
    with open(path, "w") as f:
        f.write('"""GENERATED by tools/gen_frame_data.py from the reference\'s F00 frame data -- do not edit."""\n')
        f.write("CONSTS = " + pprint.pformat(consts, width=110) + "\n\n")
        f.write("ATTACKS = " + pprint.pformat(attacks, width=110) + "\n\n")
        f.write("ACTIONS = " + pprint.pformat(actions, width=110) + "\n\n")
        f.write("# order of footsies_gym/moves.py FOOTSIES_MOVE_INDEX_TO_MOVE (moves.py:41)\n")
        f.write("MOVE_NAMES = " + pprint.pformat(names, width=110) + "\n")
        f.write("MOVE_IDS = " + pprint.pformat([a["actionID"] for a in actions], width=110) + "\n")
        f.write("MOVE_DURATIONS = " + pprint.pformat([a["frameCount"] for a in actions], width=110) + "\n")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true", help="regenerate into memory and diff against committed files")
    args = ap.parse_args()
    consts, attacks, actions = load_all()
    names = crosscheck_moves_py(actions)
    assert names == [a["actionName"] for a in actions]
    outs = {
        os.path.join(REPO, "oracle/frame_data.h"): lambda p: emit_oracle_header(consts, attacks, actions, p),
        os.path.join(REPO, "footsies_gym_b200/frame_data.py"): lambda p: emit_python(consts, attacks, actions, names, p),
    }
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    try:
        import gen_kernel_tables
        outs[os.path.join(REPO, "footsies_gym_b200/csrc/frame_tables.h")] = \
            lambda p: gen_kernel_tables.emit(consts, attacks, actions, p)
    except ImportError:
        pass
    bad = 0
    for path, fn in outs.items():
        if args.check:
            tmp = path + ".tmp"
            fn(tmp)
            same = os.path.exists(path) and open(tmp).read() == open(path).read()
            os.remove(tmp)
            print(("OK   " if same else "DIFF ") + os.path.relpath(path, REPO))
            bad += 0 if same else 1
        else:
            fn(path)
            print("wrote", os.path.relpath(path, REPO))
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
