"""ctypes binding of tests/host_emulation (the device frame logic compiled for the host).

TEST INFRASTRUCTURE ONLY: lets the parity cases exercise csrc/frame_logic.cuh against the oracle without a GPU.
The product package never imports this and has no CPU path."""
import ctypes as C
import os
import subprocess

import numpy as np
import torch

from footsies_gym_b200 import _capi

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_emulation")
LIB = os.path.join(HERE, "libkernel_logic_host.so")
CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "footsies_gym_b200", "csrc")
_lib = None


def build(force=False):
    deps = [os.path.join(HERE, "kernel_logic_host.cpp")] + [os.path.join(CSRC, f) for f in (
        "frame_logic.cuh", "tables_host.h", "state_codec.h", "frame_tables.h")]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in deps):
        return LIB
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared",
                    "-Wno-unknown-pragmas", "-o", LIB, deps[0]], check=True, capture_output=True)
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        L.he_create.restype = C.c_void_p
        L.he_create.argtypes = [C.c_int] * 7 + [C.c_longlong]
        L.he_destroy.argtypes = [C.c_void_p]
        L.he_set_skip_unactionable.argtypes = [C.c_void_p, C.c_int]
        L.he_seed.argtypes = [C.c_void_p, C.c_longlong, C.c_void_p]
        L.he_reset.argtypes = [C.c_void_p, C.c_void_p]
        L.he_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.he_get_state.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.he_set_state.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.he_set_state.restype = C.c_int
        for name, t in (("he_obs", C.c_float), ("he_reward", C.c_float), ("he_terminated", C.c_uint8),
                        ("he_info_frame", C.c_int32), ("he_info_misc", C.c_uint8), ("he_stats", C.c_uint64)):
            getattr(L, name).restype = C.POINTER(t)
            getattr(L, name).argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class HostKernelEnv:
    """The subset of the FootsiesEnv surface the parity helpers use, backed by the host-compiled kernel logic."""

    def __init__(self, num_envs=1, by_example=False, opponent=None, dense_reward=True, frame_skip=1, autoreset=True,
                 seed=0, first_env_index=0, stale_intro_input=True, device=None):
        self.num_envs = n = int(num_envs)
        self.p1_bot, self.p2_bot = bool(by_example), opponent in (None, "bot")
        L = lib()
        self.h = L.he_create(n, int(self.p1_bot), int(self.p2_bot), int(dense_reward), int(frame_skip), int(autoreset),
                             int(stale_intro_input), int(first_env_index))

        def view(fn, shape, dtype):
            return torch.from_numpy(np.ctypeslib.as_array(fn(self.h), shape=shape).view(dtype))
        self.obs = view(L.he_obs, (n, 8), np.float32)
        self.reward = view(L.he_reward, (n,), np.float32)
        self.terminated = view(L.he_terminated, (n,), np.uint8)
        self.info_frame = view(L.he_info_frame, (n,), np.int32)
        self.info_misc = view(L.he_info_misc, (n, 4), np.uint8)
        self._stats = np.ctypeslib.as_array(L.he_stats(self.h), shape=(_capi.FG_STAT_COUNT,))
        self._mask = None
        if seed is not None:
            self.seed(seed)

    def seed(self, seed, mask=None):
        m = None if mask is None else np.ascontiguousarray(np.asarray(mask), dtype=np.uint8)
        lib().he_seed(self.h, int(seed), _ptr(m))

    def reset(self, *, seed=None, options=None):
        mask = None if not options else options.get("mask")
        m = None if mask is None else np.ascontiguousarray(np.asarray(mask), dtype=np.uint8)
        if seed is not None:
            self.seed(seed, m)
        lib().he_reset(self.h, _ptr(m))

    def set_skip_unactionable(self, flag):
        lib().he_set_skip_unactionable(self.h, int(bool(flag)))

    def set_step_mask(self, mask):
        self._mask = None if mask is None else np.ascontiguousarray(np.asarray(mask), dtype=np.uint8)

    def step(self, action=None, opponent_action=None):
        def arr(a):
            if a is None:
                return None
            a = a.numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
            return np.ascontiguousarray(a, dtype=np.uint8)
        a1, a2 = arr(action), arr(opponent_action)
        if a1 is None:
            a1 = np.zeros(self.num_envs, np.uint8)
        if a2 is None:
            a2 = np.zeros(self.num_envs, np.uint8)
        lib().he_step(self.h, _ptr(a1), _ptr(a2), _ptr(self._mask))

    def get_state(self, first=0, count=None):
        count = self.num_envs - first if count is None else count
        out = np.zeros(count, dtype=_capi.env_state_dtype())
        lib().he_get_state(self.h, int(first), int(count), _ptr(out))
        return out

    def set_state(self, states, first=0):
        states = np.ascontiguousarray(states, dtype=_capi.env_state_dtype())
        if lib().he_set_state(self.h, int(first), len(states), _ptr(states)) != 0:
            raise ValueError("state not representable")

    def episode_stats(self):
        return {k: int(v) for k, v in zip(_capi.STAT_NAMES, self._stats)}

    def close(self):
        if self.h:
            lib().he_destroy(self.h)
            self.h = None


class HostKernelTorchEnv:
    """HostKernelEnv behind the FootsiesEnv surface the wrappers use (reset / step returning the observation dicts,
    set_step_mask, set_skip_unactionable), so that the batched wrappers -- including the fused FootsiesFrameSkipped path,
    which lives in the frame logic -- replay the reference wrappers' golden vectors without a GPU."""
    is_base_footsies_env = True
    opponent = None
    by_example = False
    frame_delay = 0

    def __init__(self, num_envs=1, dense_reward=True, autoreset=False, seed=0):
        from footsies_gym_b200.env import FootsiesEnv
        from footsies_gym_b200.moves import FootsiesMove
        from footsies_gym_b200.spaces import footsies_action_space, footsies_observation_space
        self.k = HostKernelEnv(num_envs=num_envs, dense_reward=dense_reward, autoreset=autoreset, seed=seed)
        self.num_envs, self.device = num_envs, torch.device("cpu")
        relevant = [m for m in FootsiesMove if m.name not in ("WIN", "DEAD")]
        self.observation_space = footsies_observation_space(len(relevant), max(m.value.duration for m in relevant))
        self.action_space = footsies_action_space()
        self.obs, self.reward, self.info_frame = self.k.obs, self.k.reward, self.k.info_frame
        self.truncated = torch.zeros(num_envs, dtype=torch.bool)
        self._obs_dict = FootsiesEnv._make_obs_dict(self.obs)

    @property
    def terminated(self):
        return self.k.terminated.bool()

    def _out(self):
        return self._obs_dict, {"frame": self.info_frame, **self._obs_dict}

    def reset(self, *, seed=None, options=None):
        self.k.reset(seed=seed, options=options)
        return self._out()

    def set_step_mask(self, mask):
        self.k.set_step_mask(None if mask is None else mask.numpy())

    def set_skip_unactionable(self, flag):
        self.k.set_skip_unactionable(flag)

    def step(self, action):
        from footsies_gym_b200.env import _as_bitmask
        self.k.step(_as_bitmask(action, self.num_envs, "cpu").numpy())
        obs, info = self._out()
        return obs, self.reward, self.terminated, self.truncated, info

    def close(self):
        self.k.close()
