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
        self.h = lib().fo_create(self.n, C.byref(self.cfg), int(first_env_index))
        self.trace = np.zeros(self.n, dtype=TRACE_DTYPE)
        if seed is not None:
            self.seed(seed)

    def __del__(self):
        if getattr(self, "h", None):
            lib().fo_destroy(self.h)
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
        lib().fo_get_trace(self.h, int(env), C.c_void_p(self.trace.ctypes.data + env * TRACE_DTYPE.itemsize))

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
