"""ctypes binding of oracle/_ref/libfootsies_ref.so: the reference's OWN C# battle code, transliterated mechanically by
tools/cs2cpp.py and driven behind the oracle's C API (oracle/ref_shim/ref_harness.cpp).

TEST INFRASTRUCTURE: the second, independent checker that pins oracle/ to the reference's source text
(tests/test_oracle_vs_ref.py).  It is (re)built only where /root/reference exists (the authoring container); on the GPU
box the prebuilt library that travelled with the snapshot is used.  The product never imports this."""
import ctypes as C
import os
import subprocess

import numpy as np

import oracle_binding as ob

ORACLE_DIR = ob.ORACLE_DIR
LIB_PATH = os.path.join(ORACLE_DIR, "_ref", "libfootsies_ref.so")
REF_SCRIPTS = "/root/reference/Assets/Script"


def available():
    return os.path.isdir(REF_SCRIPTS) or os.path.exists(LIB_PATH)


def build(force=False):
    if os.path.isdir(REF_SCRIPTS):
        cmd = ["make", "-C", ORACLE_DIR, "-f", "Makefile.ref"] + (["-B"] if force else [])
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("building oracle/_ref failed:\n" + res.stdout + res.stderr)
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("oracle/_ref/libfootsies_ref.so is missing and /root/reference is not here to generate it")
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.fo_create.restype = C.c_void_p
        L.fo_create.argtypes = [C.c_int32, C.POINTER(ob.Config), C.c_int64]
        L.fo_destroy.argtypes = [C.c_void_p]
        L.fo_seed.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
        L.fo_set_rng_tape.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32]
        L.fo_reset.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.fo_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32]
        L.fo_get_trace.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        L.fo_set_state.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32]
        L.fo_save_battle_state.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        L.fo_load_battle_state.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        L.fo_frames_simulated.restype = C.c_int64
        L.fo_frames_simulated.argtypes = [C.c_void_p]
        L.fo_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


class RefBatch(ob.OracleBatch):
    """OracleBatch's interface over the transliterated reference engine (one game object graph per battle)."""

    def __init__(self, num_envs, p1_bot=False, p2_bot=True, dense_reward=True, frame_delay=0,
                 autoreset=True, stale_intro_input=True, first_env_index=0, seed=0, threads=1):
        self._L = lib()
        self.n = int(num_envs)
        self.threads = int(threads)
        self.cfg = ob.Config(int(p1_bot), int(p2_bot), int(dense_reward), int(frame_delay), int(autoreset),
                             int(stale_intro_input))
        self.h = self._L.fo_create(self.n, C.byref(self.cfg), int(first_env_index))
        self.trace = np.zeros(self.n, dtype=ob.TRACE_DTYPE)
        if seed is not None:
            self.seed(seed)

    def __del__(self):
        if getattr(self, "h", None):
            self._L.fo_destroy(self.h)
            self.h = None

    def seed(self, seed_base, mask=None):
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        self._L.fo_seed(self.h, int(seed_base), ob._ptr(m))

    def set_rng_tape(self, env, raw):
        raw = np.ascontiguousarray(raw, dtype=np.uint32)
        self._L.fo_set_rng_tape(self.h, int(env), ob._ptr(raw), len(raw))

    def reset(self, mask=None):
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        self._L.fo_reset(self.h, ob._ptr(m), ob._ptr(self.trace))
        return self.trace

    def step(self, a1, a2=None, repeat=1):
        a1 = np.ascontiguousarray(a1, dtype=np.uint8)
        if a2 is not None:
            a2 = np.ascontiguousarray(a2, dtype=np.uint8)
        self._L.fo_step(self.h, ob._ptr(a1), ob._ptr(a2), int(repeat), ob._ptr(self.trace), self.threads)
        return self.trace

    def save_battle_state(self, env):
        raw = np.zeros(1, dtype=ob.BATTLE_STATE_DTYPE)
        self._L.fo_save_battle_state(self.h, int(env), ob._ptr(raw))
        return ob.battle_state_record_to_dict(raw[0])

    def load_battle_state(self, env, state_dict):
        raw = ob.battle_state_dict_to_record(state_dict)
        self._L.fo_load_battle_state(self.h, int(env), ob._ptr(raw))
        self._L.fo_get_trace(self.h, int(env), C.c_void_p(self.trace.ctypes.data + int(env) * ob.TRACE_DTYPE.itemsize))

    def frames_simulated(self):
        return int(self._L.fo_frames_simulated(self.h))

    def set_state(self, env, p1, p2, frame=0):
        s1 = np.zeros(1, dtype=ob.FIGHTER_DTYPE)
        s2 = np.zeros(1, dtype=ob.FIGHTER_DTYPE)
        for s, d in ((s1, p1), (s2, p2)):
            s["guard"], s["vital"], s["buffer_id"], s["reserve_id"] = 3, 1, -1, -1
            for k, v in d.items():
                s[k] = v
        self._L.fo_set_state(self.h, int(env), ob._ptr(s1), ob._ptr(s2), int(frame))
        self._L.fo_get_trace(self.h, int(env), C.c_void_p(self.trace.ctypes.data + int(env) * ob.TRACE_DTYPE.itemsize))

    def stats(self):
        out = np.zeros(len(ob.STAT_NAMES), dtype=np.int64)
        ret = C.c_double(0.0)
        self._L.fo_stats(self.h, ob._ptr(out), C.byref(ret))
        d = {k: int(v) for k, v in zip(ob.STAT_NAMES, out)}
        d["return_sum"] = ret.value
        return d
