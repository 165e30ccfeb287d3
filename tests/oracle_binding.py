"""ctypes binding of the CPU oracle (oracle/libfootsies_oracle.so).

TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs.  The product package (footsies_gym_b200) never imports this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ORACLE_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "libfootsies_oracle.so")

FIGHTER_DTYPE = np.dtype([
    ("pos_x", "<f4"), ("velocity_x", "<f4"), ("action_id", "<i4"), ("action_frame", "<i4"),
    ("hitstun", "<i4"), ("guard", "<i4"), ("vital", "<i4"), ("hit_count", "<i4"),
    ("buffer_id", "<i4"), ("reserve_id", "<i4"), ("is_input_backward", "<i4"), ("is_reserve_prox", "<i4"),
    ("shake", "<i4"), ("has_won", "<i4"), ("input0", "<i4"), ("hist_left", "<u4"), ("hist_right", "<u4"),
    ("attack_run", "<i4"),
])
TRACE_DTYPE = np.dtype([
    ("f", FIGHTER_DTYPE, (2,)), ("frame", "<i4"), ("recorded_input", "<i4", (2,)), ("events", "<i4"),
    ("battle_over", "<i4"), ("was_reset", "<i4"), ("rng_draws", "<i4"), ("rng_state", "<u4", (4,)),
    ("bot_input", "<i4", (2,)), ("obs", "<f4", (8,)), ("reward", "<f4"), ("terminated", "<i4"),
    ("info_frame", "<i4"), ("info_action", "<i4", (2,)), ("info_hitstun", "<i4", (2,)),
    ("reward_f64", "<f8"),
], align=True)

RECT_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("width", "<f4"), ("height", "<f4")])
HITBOX_DTYPE = np.dtype([("rect", RECT_DTYPE), ("proximity", "<i4"), ("attackID", "<i4")])
FULL_FIGHTER_DTYPE = np.dtype([
    ("position", "<f4", (2,)), ("velocity_x", "<f4"), ("isFaceRight", "<i4"),
    ("n_hitboxes", "<i4"), ("hitboxes", HITBOX_DTYPE, (8,)), ("n_hurtboxes", "<i4"), ("hurtboxes", RECT_DTYPE, (8,)),
    ("pushbox", RECT_DTYPE), ("vitalHealth", "<i4"), ("guardHealth", "<i4"), ("currentActionID", "<i4"),
    ("currentActionFrame", "<i4"), ("currentActionHitCount", "<i4"), ("currentHitStunFrame", "<i4"),
    ("input", "<i4", (180,)), ("inputDown", "<i4", (180,)), ("inputUp", "<i4", (180,)),
    ("isInputBackward", "<i4"), ("isReserveProximityGuard", "<i4"), ("bufferActionID", "<i4"),
    ("reserveDamageActionID", "<i4"), ("spriteShakePosition", "<i4"), ("maxSpriteShakeFrame", "<i4"),
    ("hasWon", "<i4"),
])
BATTLE_STATE_DTYPE = np.dtype([("p", FULL_FIGHTER_DTYPE, (2,)), ("roundStartTime", "<f4"), ("frameCount", "<i4")])

STAT_NAMES = ["episodes", "p1_wins", "p2_wins", "double_ko", "frames", "p1_specials", "p1_specials_neutral",
              "guard_breaks", "hits", "blocks"]


class Config(C.Structure):
    _fields_ = [("p1_bot", C.c_int32), ("p2_bot", C.c_int32), ("dense_reward", C.c_int32),
                ("frame_delay", C.c_int32), ("autoreset", C.c_int32), ("stale_intro_input", C.c_int32)]


def build(force=False):
    src = os.path.join(ORACLE_DIR, "footsies_oracle.c")
    deps = [src, os.path.join(ORACLE_DIR, "footsies_oracle.h"), os.path.join(ORACLE_DIR, "frame_data.h")]
    if (not force and os.path.exists(LIB_PATH)
            and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(d) for d in deps)):
        return LIB_PATH
    subprocess.run(["make", "-C", ORACLE_DIR, "-B"], check=True, capture_output=True)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.fo_create.restype = C.c_void_p
        L.fo_create.argtypes = [C.c_int32, C.POINTER(Config), C.c_int64]
        L.fo_destroy.argtypes = [C.c_void_p]
        L.fo_seed.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
        L.fo_set_rng_tape.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32]
        L.fo_reset.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.fo_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32]
        L.fo_set_state.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32]
        L.fo_get_trace.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        L.fo_save_battle_state.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        L.fo_load_battle_state.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        L.fo_frames_simulated.restype = C.c_int64
        L.fo_frames_simulated.argtypes = [C.c_void_p]
        L.fo_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.fo_rng_init.argtypes = [C.c_void_p, C.c_int32]
        L.fo_rng_next.restype = C.c_uint32
        L.fo_rng_next.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class OracleBatch:
    """N independent reference battles stepped on the CPU."""

    def __init__(self, num_envs, p1_bot=False, p2_bot=True, dense_reward=True, frame_delay=0,
                 autoreset=True, stale_intro_input=True, first_env_index=0, seed=0, threads=1):
        assert TRACE_DTYPE.itemsize == 264, TRACE_DTYPE.itemsize
        self.n = int(num_envs)
        self.threads = int(threads)
        self.cfg = Config(int(p1_bot), int(p2_bot), int(dense_reward), int(frame_delay), int(autoreset),
                          int(stale_intro_input))
        self._L = lib()                       # kept so that __del__ still works during interpreter shutdown
        self.h = self._L.fo_create(self.n, C.byref(self.cfg), int(first_env_index))
        self.trace = np.zeros(self.n, dtype=TRACE_DTYPE)
        if seed is not None:
            self.seed(seed)

    def __del__(self):
        if getattr(self, "h", None):
            self._L.fo_destroy(self.h)
            self.h = None

    def seed(self, seed_base, mask=None):
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        lib().fo_seed(self.h, int(seed_base), _ptr(m))

    def set_rng_tape(self, env, raw):
        raw = np.ascontiguousarray(raw, dtype=np.uint32)
        lib().fo_set_rng_tape(self.h, int(env), _ptr(raw), len(raw))

    def reset(self, mask=None):
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        lib().fo_reset(self.h, _ptr(m), _ptr(self.trace))
        return self.trace

    def step(self, a1, a2=None, repeat=1):
        a1 = np.ascontiguousarray(a1, dtype=np.uint8)
        assert a1.shape == (self.n,)
        if a2 is not None:
            a2 = np.ascontiguousarray(a2, dtype=np.uint8)
            assert a2.shape == (self.n,)
        lib().fo_step(self.h, _ptr(a1), _ptr(a2), int(repeat), _ptr(self.trace), self.threads)
        return self.trace

    def set_state(self, env, p1, p2, frame=0):
        s1 = np.zeros(1, dtype=FIGHTER_DTYPE)
        s2 = np.zeros(1, dtype=FIGHTER_DTYPE)
        for s, d in ((s1, p1), (s2, p2)):
            s["guard"] = 3
            s["vital"] = 1
            s["buffer_id"] = -1
            s["reserve_id"] = -1
            for k, v in d.items():
                s[k] = v
        lib().fo_set_state(self.h, int(env), _ptr(s1), _ptr(s2), int(frame))
        lib().fo_get_trace(self.h, int(env), C.c_void_p(self.trace.ctypes.data + int(env) * TRACE_DTYPE.itemsize))

    # ---- full battle state in the reference's save / load schema (BattleCore.SaveState / LoadState) ----
    def save_battle_state(self, env):
        """-> dict in the JSON layout of BattleState.cs / FighterState.cs (what JsonUtility.ToJson would emit)."""
        raw = np.zeros(1, dtype=BATTLE_STATE_DTYPE)
        lib().fo_save_battle_state(self.h, int(env), _ptr(raw))
        return battle_state_record_to_dict(raw[0])

    def load_battle_state(self, env, state_dict):
        raw = battle_state_dict_to_record(state_dict)
        lib().fo_load_battle_state(self.h, int(env), _ptr(raw))
        lib().fo_get_trace(self.h, int(env), C.c_void_p(self.trace.ctypes.data + int(env) * TRACE_DTYPE.itemsize))

    def frames_simulated(self):
        return int(lib().fo_frames_simulated(self.h))

    def stats(self):
        out = np.zeros(len(STAT_NAMES), dtype=np.int64)
        ret = C.c_double(0.0)
        lib().fo_stats(self.h, _ptr(out), C.byref(ret))
        d = {k: int(v) for k, v in zip(STAT_NAMES, out)}
        d["return_sum"] = ret.value
        return d


def rng_stream(seed, n):
    s = (C.c_uint32 * 4)()
    lib().fo_rng_init(s, int(seed))
    return [int(lib().fo_rng_next(s)) for _ in range(n)]


def _rect_dict(r):
    return {"x": float(r["x"]), "y": float(r["y"]), "width": float(r["width"]), "height": float(r["height"])}


def battle_state_record_to_dict(rec):
    out = {}
    for name, f in (("p1State", rec["p"][0]), ("p2State", rec["p"][1])):
        out[name] = {
            "position": [float(f["position"][0]), float(f["position"][1])], "velocity_x": float(f["velocity_x"]),
            "isFaceRight": bool(f["isFaceRight"]),
            "hitboxes": [{"rect": _rect_dict(f["hitboxes"][k]["rect"]), "proximity": bool(f["hitboxes"][k]["proximity"]),
                          "attackID": int(f["hitboxes"][k]["attackID"])} for k in range(int(f["n_hitboxes"]))],
            "hurtboxes": [_rect_dict(f["hurtboxes"][k]) for k in range(int(f["n_hurtboxes"]))],
            "pushbox": _rect_dict(f["pushbox"]),
            "vitalHealth": int(f["vitalHealth"]), "guardHealth": int(f["guardHealth"]),
            "currentActionID": int(f["currentActionID"]), "currentActionFrame": int(f["currentActionFrame"]),
            "currentActionHitCount": int(f["currentActionHitCount"]), "currentHitStunFrame": int(f["currentHitStunFrame"]),
            "input": [int(v) for v in f["input"]], "inputDown": [int(v) for v in f["inputDown"]],
            "inputUp": [int(v) for v in f["inputUp"]],
            "isInputBackward": bool(f["isInputBackward"]), "isReserveProximityGuard": bool(f["isReserveProximityGuard"]),
            "bufferActionID": int(f["bufferActionID"]), "reserveDamageActionID": int(f["reserveDamageActionID"]),
            "spriteShakePosition": int(f["spriteShakePosition"]), "maxSpriteShakeFrame": int(f["maxSpriteShakeFrame"]),
            "hasWon": bool(f["hasWon"]),
        }
    out["roundStartTime"] = float(rec["roundStartTime"])
    out["frameCount"] = int(rec["frameCount"])
    return out


def battle_state_dict_to_record(d):
    raw = np.zeros(1, dtype=BATTLE_STATE_DTYPE)
    for i, name in enumerate(("p1State", "p2State")):
        s, f = d[name], raw[0]["p"][i]
        f["position"] = s["position"]
        f["velocity_x"] = s["velocity_x"]
        f["isFaceRight"] = int(s["isFaceRight"])
        f["n_hitboxes"] = len(s["hitboxes"])
        for k, h in enumerate(s["hitboxes"]):
            f["hitboxes"][k]["rect"] = (h["rect"]["x"], h["rect"]["y"], h["rect"]["width"], h["rect"]["height"])
            f["hitboxes"][k]["proximity"] = int(h["proximity"])
            f["hitboxes"][k]["attackID"] = h["attackID"]
        f["n_hurtboxes"] = len(s["hurtboxes"])
        for k, h in enumerate(s["hurtboxes"]):
            f["hurtboxes"][k] = (h["x"], h["y"], h["width"], h["height"])
        p = s["pushbox"]
        f["pushbox"] = (p["x"], p["y"], p["width"], p["height"])
        for key in ("vitalHealth", "guardHealth", "currentActionID", "currentActionFrame", "currentActionHitCount",
                    "currentHitStunFrame", "bufferActionID", "reserveDamageActionID", "spriteShakePosition",
                    "maxSpriteShakeFrame"):
            f[key] = s[key]
        for key in ("input", "inputDown", "inputUp"):
            arr = list(s[key])[:180]
            f[key][:len(arr)] = arr
        for key in ("isInputBackward", "isReserveProximityGuard", "hasWon"):
            f[key] = int(s[key])
    raw[0]["roundStartTime"] = d["roundStartTime"]
    raw[0]["frameCount"] = d["frameCount"]
    return raw
