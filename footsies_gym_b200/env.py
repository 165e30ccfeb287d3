"""Batched, B200-native drop-in for the reference's FootsiesEnv (footsies-gym/footsies_gym/envs/footsies.py).

Same API surface -- reset / step / hard_reset / close / set_opponent / observation_space / action_space /
reward_range / most_recent_observation / most_recent_info -- with a leading `num_envs` dimension: every
call steps N independent battles on one GPU through libfootsies_b200.so (include/footsies_b200.h).  No game
binary, no sockets, no CPU fallback.

Differences from the single-process reference that a caller must know:
  * observations / rewards / flags are torch tensors on the device, written IN PLACE by the kernel: the
    tensors returned by step() are the same objects every call (clone what you need to keep);
  * obs["guard"], obs["move"], obs["move_frame"], obs["position"] are float32 views [N, 2] of one [N, 8]
    tensor (`env.obs`), which a policy network can consume directly;
  * actions are a uint8 bitmask per env (Left=1, Right=2, Attack=4: the game's InputDefine,
    InputData.cs:8-14, i.e. what wrappers/action_comb_disc.py produces) or an [N, 3] 0/1 tensor in the
    reference's MultiBinary(3) order (left, right, attack);
  * finished battles restart by themselves like the game does (BattleCore.cs:176-180): with
    autoreset=True the step() after a terminal one returns the first observation of the next battle
    (frame -1, reward 0, terminated False) and ignores the action, exactly the state the reference's
    reset() would have read from the socket.
"""
import ctypes as C
from typing import Callable, Optional, Union

import numpy as np
import torch

from . import _capi
from . import frame_data as _fd
from .moves import FootsiesMove
from .spaces import footsies_action_space, footsies_observation_space


class FootsiesGameClosedError(RuntimeError):
    """Kept for API compatibility with footsies_gym.envs.exceptions (never raised: there is no game process)."""


def _as_bitmask(action, n, device):
    """Accepts a uint8 bitmask [N], an [N, 3] (left, right, attack) array, or one 3-tuple for N == 1."""
    if isinstance(action, torch.Tensor):
        t = action
    else:
        t = torch.as_tensor(np.asarray(action))
    if t.dim() == 1 and t.shape[0] == 3 and n == 1 and t.dtype in (torch.bool,):
        t = t.reshape(1, 3)
    if t.dim() == 2:
        if t.shape != (n, 3):
            raise ValueError(f"action must have shape ({n},) or ({n}, 3), got {tuple(t.shape)}")
        t = t.to(torch.uint8)
        t = t[:, 0] | (t[:, 1] << 1) | (t[:, 2] << 2)
    elif t.dim() == 1:
        if t.shape[0] == 3 and n == 1:
            t = t.to(torch.uint8)
            t = (t[0] | (t[1] << 1) | (t[2] << 2)).reshape(1)
        elif t.shape[0] != n:
            raise ValueError(f"action must have shape ({n},) or ({n}, 3), got {tuple(t.shape)}")
    elif t.dim() == 0 and n == 1:
        t = t.reshape(1)
    else:
        raise ValueError(f"unsupported action shape {tuple(t.shape)}")
    # a device -> host copy into pageable memory is only safe as a blocking copy (the caller reads it right away)
    to_cpu = torch.device(device).type == "cpu"
    return t.to(device=device, dtype=torch.uint8, non_blocking=not to_cpu)


class FootsiesEnv:
    metadata = {"render_modes": [], "render_fps": 50}

    def __init__(
        self,
        num_envs: int = 1,
        device: Union[str, torch.device, None] = None,
        frame_delay: int = 0,
        render_mode: Optional[str] = None,
        by_example: bool = False,
        opponent: Union[Callable, str, None] = None,
        vs_player: bool = False,
        dense_reward: bool = True,
        frame_skip: int = 1,
        autoreset: bool = True,
        seed: Optional[int] = 0,
        first_env_index: int = 0,
        stale_intro_input: bool = True,
        skip_unactionable: bool = False,
        **reference_kwargs,
    ):
        """
        Parameters (reference names keep their meaning, footsies.py:34-97)
        ----------
        num_envs: battles stepped per call on this GPU
        device: CUDA device (default: current device)
        frame_delay: observations / info are delayed by this many env steps (reward and termination are not): frames when
                     frame_skip is 1, as in the reference (footsies.py:129-131, 533-535); with K fused frames per step the queue
                     still advances once per step() call, i.e. the delay is frame_delay x K frames
        by_example: P1 is driven by the in-game bot, actions passed to step() are ignored
        opponent: None or "bot" -> in-game BattleAI (reference default); a callable `(obs, info) -> actions`
            -> custom policy queried every step like the reference's `opponent`; "self_play" / "remote"
            -> P2 actions are passed to step(action, opponent_action)
        dense_reward: +-0.3 per guard point with terminal compensation, else sparse +-1
        frame_skip: K >= 1 frames fused per step() with the action repeated; rewards are summed
        autoreset: see module docstring
        seed: per-env bot RNG = InitState(seed + first_env_index + i); None leaves the RNG planes zeroed
        first_env_index: global index of env 0 (multi-GPU sharding keeps results independent of the split)
        skip_unactionable: FootsiesFrameSkipped (wrappers/frame_skip.py:46-80) fused into step(): an env whose new
            observation P1 cannot act on keeps stepping with P1's no-op action inside the same kernel launch, rewards
            summed (the FootsiesFrameSkipped wrapper switches this on; see set_skip_unactionable)
        reference_kwargs: game_path, game_address, game_port, fast_forward, sync_mode, ... are accepted
            and ignored (there is no game process to configure)
        """
        if render_mode is not None:
            raise ValueError("render_mode is not supported: this simulator is headless")
        if vs_player:
            raise ValueError("vs_player is not supported: there is no human input device")
        if opponent is not None and not callable(opponent) and opponent not in ("bot", "self_play", "remote"):
            raise ValueError("opponent must be None, 'bot', 'self_play', 'remote' or a callable")
        unknown = set(reference_kwargs) - {
            "game_path", "game_address", "game_port", "skip_instancing", "fast_forward", "fast_forward_speed",
            "sync_mode", "remote_control_port", "opponent_port", "log_file", "log_file_overwrite"}
        if unknown:
            raise TypeError(f"unexpected keyword arguments: {sorted(unknown)}")
        if num_envs < 1:
            raise ValueError("num_envs must be >= 1")
        if frame_delay < 0:
            raise ValueError("frame_delay must be >= 0")
        if not torch.cuda.is_available():
            raise RuntimeError("footsies_gym_b200 needs a CUDA device (sm_100a); there is no CPU fallback")

        self.num_envs = int(num_envs)
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        if self.device.type != "cuda":
            raise RuntimeError("footsies_gym_b200 runs on CUDA devices only")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.frame_delay = int(frame_delay)
        self.by_example = bool(by_example)
        self.opponent = opponent if callable(opponent) else None
        self._opponent_mode = "bot" if opponent in (None, "bot") else "remote"
        self.dense_reward = bool(dense_reward)
        self.frame_skip = int(frame_skip)
        self.autoreset = bool(autoreset)
        self.first_env_index = int(first_env_index)
        self._reward_table = None
        self._load_fix = {}          # env index -> {"cum", "guard"}: battles whose reward is host-tracked after a load
        self.stale_intro_input = bool(stale_intro_input)
        self.skip_unactionable = bool(skip_unactionable)
        if self.skip_unactionable and self.by_example:
            raise ValueError("skip_unactionable needs an agent-controlled P1 (by_example=False)")
        self.render_mode = None

        relevant = [m for m in FootsiesMove if m.name not in ("WIN", "DEAD")]
        self.observation_space = footsies_observation_space(len(relevant), max(m.value.duration for m in relevant))
        self.action_space = footsies_action_space()
        self.reward_range = (-1, 1)
        # the spaces above describe ONE battle, as in the reference; the names a gymnasium VectorEnv user looks for
        self.single_observation_space, self.single_action_space = self.observation_space, self.action_space

        self._lib = _capi.load()
        self._handle = None
        self._step_mask = None
        self._allocate()
        self._create_handle()
        self._seed_value = seed
        if seed is not None:
            self.seed(seed)
        self.has_reset = False
        self._most_recent_observation = None
        self._most_recent_info = None

    # ------------------------------------------------------------------ buffers / handle
    def _allocate(self):
        n, dev = self.num_envs, self.device
        self.state = torch.zeros((_capi.FG_STATE_PLANES, n, 4), dtype=torch.int32, device=dev)
        self.stats = torch.zeros(_capi.FG_STAT_COUNT, dtype=torch.int64, device=dev)
        self.actions_p1 = torch.zeros(n, dtype=torch.uint8, device=dev)
        self.actions_p2 = torch.zeros(n, dtype=torch.uint8, device=dev)
        self.obs = torch.zeros((n, 8), dtype=torch.float32, device=dev)
        self.reward = torch.zeros(n, dtype=torch.float32, device=dev)
        self.terminated = torch.zeros(n, dtype=torch.bool, device=dev)
        self.truncated = torch.zeros(n, dtype=torch.bool, device=dev)   # the reference never truncates (footsies.py:570)
        self.info_frame = torch.zeros(n, dtype=torch.int32, device=dev)
        self.info_misc = torch.zeros((n, 4), dtype=torch.uint8, device=dev)
        self._obs_dict = self._make_obs_dict(self.obs)
        self._info_dict = self._make_info_dict(self.info_frame, self.info_misc, self._obs_dict)
        if self.frame_delay > 0:
            d = self.frame_delay + 1
            self._ring_obs = torch.zeros((d, n, 8), dtype=torch.float32, device=dev)
            self._ring_frame = torch.zeros((d, n), dtype=torch.int32, device=dev)
            self._ring_misc = torch.zeros((d, n, 4), dtype=torch.uint8, device=dev)
            self._ring_pos = 0
            self._delayed_obs = torch.zeros((n, 8), dtype=torch.float32, device=dev)
            self._delayed_frame = torch.zeros(n, dtype=torch.int32, device=dev)
            self._delayed_misc = torch.zeros((n, 4), dtype=torch.uint8, device=dev)
            self._delayed_obs_dict = self._make_obs_dict(self._delayed_obs)
            self._delayed_info_dict = self._make_info_dict(self._delayed_frame, self._delayed_misc, self._delayed_obs_dict)
        # pinned host mirrors for the host-buffer (reference-facing) path, created on first use
        self._host = None

    @staticmethod
    def _make_obs_dict(obs):
        return {"guard": obs[:, 0:2], "move": obs[:, 2:4], "move_frame": obs[:, 4:6], "position": obs[:, 6:8]}

    @staticmethod
    def _make_info_dict(frame, misc, obs_dict):
        d = {"frame": frame, "p1_action": misc[:, 0], "p2_action": misc[:, 1],
             "p1_hitstun": misc[:, 2], "p2_hitstun": misc[:, 3]}
        d.update(obs_dict)   # the reference copies the observation into info (footsies.py:378-379)
        return d

    def _config(self):
        return _capi.FgConfig(
            struct_size=C.sizeof(_capi.FgConfig), num_envs=self.num_envs, device=self.device.index,
            p1_bot=int(self.by_example), p2_bot=int(self._opponent_mode == "bot"),
            dense_reward=int(self.dense_reward), frame_skip=self.frame_skip, autoreset=int(self.autoreset),
            stale_intro_input=int(self.stale_intro_input), skip_unactionable=int(self.skip_unactionable),
            first_env_index=self.first_env_index)

    def _create_handle(self):
        if self._handle is not None:
            self._lib.fg_destroy(self._handle)
            self._handle = None
        h = C.c_void_p()
        cfg = self._config()
        _capi.check(self._lib.fg_create(C.byref(cfg), C.byref(h)))
        self._handle = h
        cfg2 = self._config()
        self._bind()
        self.algorithmic_bytes_per_env_step = int(self._lib.fg_algorithmic_bytes_per_env_step(C.byref(cfg2)))

    def _bind(self):
        b = _capi.FgBuffers()
        b.struct_size = C.sizeof(_capi.FgBuffers)
        for k in range(_capi.FG_STATE_PLANES):
            b.state[k] = self.state[k].data_ptr()
        b.stats = self.stats.data_ptr()
        b.actions_p1 = self.actions_p1.data_ptr()
        b.actions_p2 = self.actions_p2.data_ptr()
        b.obs = self.obs.data_ptr()
        b.reward = self.reward.data_ptr()
        b.terminated = self.terminated.data_ptr()
        b.info_frame = self.info_frame.data_ptr()
        b.info_misc = self.info_misc.data_ptr()
        b.step_mask = None if self._step_mask is None else self._step_mask.data_ptr()
        _capi.check(self._lib.fg_bind(self._handle, C.byref(b)))

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ------------------------------------------------------------------ reference API
    def seed(self, seed: int, mask: Optional[torch.Tensor] = None):
        """Remote-control SEED (footsies.py:454-456): env i gets Random.InitState(seed + first_env_index + i)."""
        m = None if mask is None else mask.to(device=self.device, dtype=torch.uint8).contiguous()
        _capi.check(self._lib.fg_seed(self._handle, int(seed), None if m is None else C.c_void_p(m.data_ptr()),
                                      self._stream()))

    def reset(self, *, seed: Optional[int] = None, options: Optional[dict] = None):
        """Reset every env (or options={'mask': bool tensor [N]}); returns (obs, info) of frame -1 (footsies.py:482-515)."""
        mask = None if not options else options.get("mask")
        if seed is not None:
            self.seed(seed, mask)
        return self.hard_reset(mask)

    def hard_reset(self, mask: Optional[torch.Tensor] = None):
        """Force the selected envs (default all) back to the start of a battle (README.md:87, RESET command)."""
        m = None if mask is None else mask.to(device=self.device, dtype=torch.uint8).contiguous()
        _capi.check(self._lib.fg_reset(self._handle, None if m is None else C.c_void_p(m.data_ptr()), self._stream()))
        self.has_reset = True
        if self.frame_delay > 0:
            sel = slice(None) if m is None else m.bool()
            self._ring_obs[:, sel] = self.obs[sel]          # queue pre-filled with the first state (footsies.py:502-504)
            self._ring_frame[:, sel] = self.info_frame[sel]
            self._ring_misc[:, sel] = self.info_misc[sel]
            self._delayed_obs[sel] = self.obs[sel]
            self._delayed_frame[sel] = self.info_frame[sel]
            self._delayed_misc[sel] = self.info_misc[sel]
        return self._finish_obs()

    def step(self, action=None, opponent_action=None):
        """One FootsiesEnv.step for all envs: returns (obs, reward, terminated, truncated, info) (footsies.py:518-570)."""
        if not self.has_reset:
            raise RuntimeError("call reset() before step()")
        if not self.by_example:
            if action is None:
                raise ValueError("action is required unless by_example=True")
            self.actions_p1.copy_(_as_bitmask(action, self.num_envs, self.device), non_blocking=True)
        if self._opponent_mode != "bot":
            if opponent_action is None:
                if self.opponent is None:
                    raise ValueError("opponent_action is required when the opponent is not the in-game bot")
                opponent_action = self.opponent(self._most_recent_observation, self._most_recent_info)
            self.actions_p2.copy_(_as_bitmask(opponent_action, self.num_envs, self.device), non_blocking=True)
        _capi.check(self._lib.fg_step(self._handle, self._stream()))
        if self._load_fix:
            self._apply_load_fix()
        if self.frame_delay > 0:
            self._advance_delay_ring()
        obs, info = self._finish_obs()
        return obs, self.reward, self.terminated, self.truncated, info

    def bind_actions(self, actions_p1: Optional[torch.Tensor] = None, actions_p2: Optional[torch.Tensor] = None):
        """Zero-copy action input: make the kernel read its actions straight from the given device tensors
        (uint8 [N] bitmasks, e.g. a policy's output buffer) from now on; then call step_bound()."""
        for t in (actions_p1, actions_p2):
            if t is not None and (t.dtype != torch.uint8 or t.shape != (self.num_envs,) or t.device != self.device
                                  or not t.is_contiguous()):
                raise ValueError("bound action tensors must be contiguous uint8 [num_envs] on the env's device")
        if actions_p1 is not None:
            self.actions_p1 = actions_p1
        if actions_p2 is not None:
            self.actions_p2 = actions_p2
        self._bind()

    def bind_outputs(self, obs: Optional[torch.Tensor] = None, reward: Optional[torch.Tensor] = None,
                     terminated: Optional[torch.Tensor] = None):
        """Zero-copy outputs: make the kernel write its observations / rewards / termination flags straight into the
        given device tensors from now on (e.g. slot t of a rollout buffer): obs float32 [N, 8], reward float32 [N],
        terminated bool or uint8 [N], all contiguous.  `env.obs` and the observation dicts follow."""
        n, dev = self.num_envs, self.device
        for t, shape, dts in ((obs, (n, 8), (torch.float32,)), (reward, (n,), (torch.float32,)),
                              (terminated, (n,), (torch.bool, torch.uint8))):
            if t is not None and (t.dtype not in dts or tuple(t.shape) != shape or t.device != dev or not t.is_contiguous()):
                raise ValueError(f"bound output tensors must be contiguous {dts} {shape} on the env's device")
        if obs is not None:
            self.obs = obs
            self._obs_dict = self._make_obs_dict(self.obs)
            self._info_dict = self._make_info_dict(self.info_frame, self.info_misc, self._obs_dict)
        if reward is not None:
            self.reward = reward
        if terminated is not None:
            self.terminated = terminated
        self._bind()

    def step_bound(self):
        """step() without any action copy: the kernel reads the tensors given to bind_actions()."""
        if not self.has_reset:
            raise RuntimeError("call reset() before step()")
        _capi.check(self._lib.fg_step(self._handle, self._stream()))
        if self._load_fix:
            self._apply_load_fix()
        if self.frame_delay > 0:
            self._advance_delay_ring()
        obs, info = self._finish_obs()
        return obs, self.reward, self.terminated, self.truncated, info

    def set_skip_unactionable(self, flag: bool):
        """Switch the fused FootsiesFrameSkipped stepping on or off (fg_config.skip_unactionable).  The battle state
        lives in tensors this object owns, so only the library handle is re-created."""
        flag = bool(flag)
        if flag and self.by_example:
            raise ValueError("skip_unactionable needs an agent-controlled P1 (by_example=False)")
        if flag != self.skip_unactionable:
            self.skip_unactionable = flag
            torch.cuda.synchronize(self.device)
            self._create_handle()

    def set_step_mask(self, mask: Optional[torch.Tensor]):
        """Only envs with mask[i] != 0 are advanced by the following step() calls (None = all); the others keep
        their state and their last outputs."""
        if mask is None:
            self._step_mask = None
        else:
            m = mask.to(device=self.device, dtype=torch.uint8).contiguous()
            if m.shape != (self.num_envs,):
                raise ValueError("step mask must have shape (num_envs,)")
            self._step_mask = m
        self._bind()

    def _advance_delay_ring(self):
        # footsies.py:533-535: append the newest state, pop the oldest (queue length frame_delay + 1); an env that was just
        # auto-reset restarts with a queue full of its first state (footsies.py:502-504).  One kernel (fg_delay_ring_step).
        # The DEAD -> STAND remap is applied to the delayed state when it is emitted (footsies.py:538-552); the step kernel
        # already applied it to the observation it wrote, so the ring holds remapped states.
        ptr = lambda t: C.c_void_p(t.data_ptr())   # noqa: E731
        _capi.check(self._lib.fg_delay_ring_step(
            self._handle, self.frame_delay + 1, self._ring_pos, ptr(self._ring_obs), ptr(self._ring_frame), ptr(self._ring_misc),
            ptr(self._delayed_obs), ptr(self._delayed_frame), ptr(self._delayed_misc), self._stream()))
        self._ring_pos = (self._ring_pos + 1) % (self.frame_delay + 1)

    def _finish_obs(self):
        if self.frame_delay > 0:
            obs, info = self._delayed_obs_dict, self._delayed_info_dict
        else:
            obs, info = self._obs_dict, self._info_dict
        self._most_recent_observation, self._most_recent_info = obs, info
        return obs, info

    def close(self):
        if getattr(self, "_handle", None) is not None:
            self._lib.fg_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def set_opponent(self, opponent: Optional[Callable]):
        """Switch P2 between a custom policy and the in-game bot (footsies.py:458-480, P2_BOT command).
        Returns True: like the reference recommends, call reset() afterwards."""
        self.opponent = opponent
        mode = "bot" if opponent is None else "remote"
        if mode != self._opponent_mode:
            self._opponent_mode = mode
            torch.cuda.synchronize(self.device)
            self._create_handle()
        return True

    @property
    def most_recent_observation(self):
        return self._most_recent_observation

    @property
    def most_recent_info(self):
        return self._most_recent_info

    # ------------------------------------------------------------------ host-buffer (reference-facing) path
    def _host_buffers(self):
        """Pinned host tensors of the host-buffer path, in the compact layout of fg_host_outputs (27 bytes per battle:
        the call is PCIe-bound, so the integer-valued observation fields travel as bytes)."""
        if self._host is None:
            n = self.num_envs
            # one pinned block from fg_host_alloc: its pages sit on the NUMA node of this GPU
            blk = _capi.HostBlock(self.device.index or 0, n * (1 + 1 + 8 + 6 + 4 + 1 + 4 + 4 + 16) + 64 * 16)
            hb = dict(
                block=blk, a1=blk.take((n,), torch.uint8), a2=blk.take((n,), torch.uint8),
                position=blk.take((n, 2), torch.float32), obs_u8=blk.take((n, 6), torch.uint8),
                reward=blk.take((n,), torch.float32), terminated=blk.take((n,), torch.bool),
                info_frame=blk.take((n,), torch.int32), info_misc=blk.take((n, 4), torch.uint8),
                packed=blk.take((n, 4), torch.int32),
                truncated=torch.zeros(n, dtype=torch.bool))     # the reference never truncates (footsies.py:570)
            out = _capi.FgHostOutputs(struct_size=C.sizeof(_capi.FgHostOutputs), reserved0=0)
            for k in ("position", "obs_u8", "reward", "terminated", "info_frame", "info_misc"):
                setattr(out, k, hb[k].data_ptr())
            hb["out"] = out
            u8 = hb["obs_u8"]
            hb["obs"] = {"guard": u8[:, 0:2], "move": u8[:, 2:4], "move_frame": u8[:, 4:6], "position": hb["position"]}
            hb["info"] = self._make_info_dict(hb["info_frame"], hb["info_misc"], hb["obs"])
            self._host = hb
        return self._host

    def _deliver_delayed_to_host(self):
        """frame_delay > 0: the delayed observation / info lives in device tensors maintained by step() (the delay ring,
        footsies.py:129-131, 533-535); pack it into the compact host layout on the device and copy it down.  Not the
        fast path (a handful of torch ops per step)."""
        hb = self._host_buffers()
        obs, info = self._finish_obs()
        u8 = torch.cat([obs["guard"], obs["move"], obs["move_frame"]], dim=1).to(torch.uint8)
        hb["obs_u8"].copy_(u8, non_blocking=True)
        hb["position"].copy_(obs["position"], non_blocking=True)
        hb["info_frame"].copy_(info["frame"], non_blocking=True)
        hb["info_misc"].copy_(torch.stack([info[k] for k in ("p1_action", "p2_action", "p1_hitstun", "p2_hitstun")], dim=1),
                              non_blocking=True)
        hb["reward"].copy_(self.reward, non_blocking=True)
        hb["terminated"].copy_(self.terminated.to(torch.bool), non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()

    def reset_host(self, *, seed: Optional[int] = None, mask=None):
        """reset() delivering (obs, info) to pinned host tensors (mask: host bool / uint8 [N] or None = all)."""
        hb = self._host_buffers()
        if self.frame_delay > 0:
            self.reset(seed=seed, options=None if mask is None else {"mask": torch.as_tensor(mask)})
            self._deliver_delayed_to_host()
            return hb["obs"], hb["info"]
        m = None
        if mask is not None:
            m = torch.as_tensor(mask).to(device="cpu", dtype=torch.uint8).contiguous()
            if m.numel() != self.num_envs:
                raise ValueError("mask must have num_envs entries")
        if seed is not None:
            self.seed(seed, None if m is None else m.to(self.device))
        _capi.check(self._lib.fg_reset_host_compact(self._handle, None if m is None else C.c_void_p(m.data_ptr()),
                                                    C.byref(hb["out"]), self._stream()))
        self.has_reset = True
        return hb["obs"], hb["info"]

    def step_host(self, action=None, opponent_action=None):
        """step() for callers that live on the CPU like the reference's agents: actions are read from host memory and
        obs / reward / terminated / info are delivered to pinned host tensors by one C call (fg_step_host_compact:
        H2D copies, step kernel, pack kernel, D2H copies, pipelined in slices for large batches).  The observation
        fields keep their natural types: guard, move, move_frame uint8 [N, 2], position float32 [N, 2] (the reference
        returns Python ints and floats, footsies.py:362-367).  Returns host tensors that are reused between calls."""
        if not self.has_reset:
            raise RuntimeError("call reset() before step()")
        hb = self._host_buffers()
        if self.frame_delay > 0:
            self.step(action, opponent_action)
            self._deliver_delayed_to_host()
            return hb["obs"], hb["reward"], hb["terminated"], hb["truncated"], hb["info"]
        p1, p2 = self._host_action_ptrs(action, opponent_action, hb)
        _capi.check(self._lib.fg_step_host_compact(self._handle, p1, p2, C.byref(hb["out"]), self._stream()))
        if self._load_fix:
            self._apply_load_fix(host_reward=hb["reward"])
        return hb["obs"], hb["reward"], hb["terminated"], hb["truncated"], hb["info"]

    def host_io_bytes_per_step(self, packed: bool = False):
        """(host->device, device->host) bytes moved by one step_host (or step_host_packed) call."""
        n = self.num_envs
        h2d = n * ((0 if self.by_example else 1) + (0 if self._opponent_mode == "bot" else 1))
        d2h = n * (16 if packed else 8 + 6 + 4 + 1 + 4 + 4)
        return h2d, d2h

    # ---- packed host layout: one 16-byte record per battle, one device->host copy per slice (fg_step_host_packed) ----
    def packed_reward_table(self) -> torch.Tensor:
        """float32 [128]: the reward values a packed record's 7-bit reward index stands for."""
        if self._reward_table is None:
            tab = np.zeros(_capi.FG_PACKED_REWARD_TABLE_SIZE, dtype=np.float32)
            cnt = C.c_int32(0)
            _capi.check(self._lib.fg_packed_reward_table(self._handle, C.c_void_p(tab.ctypes.data), C.byref(cnt)))
            self._reward_table = torch.from_numpy(tab)
        return self._reward_table

    def _host_action_ptrs(self, action, opponent_action, hb):
        p1 = p2 = None
        if not self.by_example:
            a = _as_bitmask(action, self.num_envs, "cpu")
            if a.data_ptr() != hb["a1"].data_ptr():
                if a.is_pinned() and a.is_contiguous():
                    p1 = C.c_void_p(a.data_ptr())
                else:
                    hb["a1"].copy_(a)
            if p1 is None:
                p1 = C.c_void_p(hb["a1"].data_ptr())
        if self._opponent_mode != "bot":
            if opponent_action is None:
                if self.opponent is None:
                    raise ValueError("opponent_action is required when the opponent is not the in-game bot")
                opponent_action = self.opponent(self._most_recent_observation, self._most_recent_info)
            a = _as_bitmask(opponent_action, self.num_envs, "cpu")
            if a.data_ptr() != hb["a2"].data_ptr():
                if a.is_pinned() and a.is_contiguous():
                    p2 = C.c_void_p(a.data_ptr())
                else:
                    hb["a2"].copy_(a)
            if p2 is None:
                p2 = C.c_void_p(hb["a2"].data_ptr())
        return p1, p2

    def step_host_packed(self, action=None, opponent_action=None) -> torch.Tensor:
        """step() for host callers in the densest lossless form: returns the pinned int32 [N, 4] tensor of
        fg_packed_result records (position p1, position p2 as float32 bits | w0 | w1, see include/footsies_b200.h);
        `decode_packed` turns (a slice of) it into the usual (obs, reward, terminated, info).  16 bytes per battle cross
        the host link instead of 27.  Needs frame_skip = 1, no fused frame skipping, no frame_delay."""
        if not self.has_reset:
            raise RuntimeError("call reset() before step()")
        if self.frame_delay > 0:
            raise ValueError("the packed host layout does not carry a frame_delay queue; use step_host")
        hb = self._host_buffers()
        p1, p2 = self._host_action_ptrs(action, opponent_action, hb)
        _capi.check(self._lib.fg_step_host_packed(self._handle, p1, p2, C.c_void_p(hb["packed"].data_ptr()), self._stream()))
        return hb["packed"]

    def reset_host_packed(self, *, seed: Optional[int] = None, mask=None) -> torch.Tensor:
        hb = self._host_buffers()
        m = None
        if mask is not None:
            m = torch.as_tensor(mask).to(device="cpu", dtype=torch.uint8).contiguous()
            if m.numel() != self.num_envs:
                raise ValueError("mask must have num_envs entries")
        if seed is not None:
            self.seed(seed, None if m is None else m.to(self.device))
        _capi.check(self._lib.fg_reset_host_packed(self._handle, None if m is None else C.c_void_p(m.data_ptr()),
                                                   C.c_void_p(hb["packed"].data_ptr()), self._stream()))
        self.has_reset = True
        return hb["packed"]

    def decode_packed(self, packed: torch.Tensor):
        """(obs, reward, terminated, truncated, info) from packed records [M, 4] int32 (any slice of what step_host_packed
        returned), with the dtypes of step_host: uint8 observation fields, float32 position / reward."""
        w0, w1 = packed[:, 2], packed[:, 3]
        pos = packed[:, 0:2].contiguous().view(torch.float32)

        def bits(w, shift, n):
            return ((w >> shift) & ((1 << n) - 1)).to(torch.uint8)
        obs = {"guard": torch.stack([bits(w0, 0, 2), bits(w0, 2, 2)], 1), "move": torch.stack([bits(w0, 4, 4), bits(w0, 8, 4)], 1),
               "move_frame": torch.stack([bits(w0, 12, 6), bits(w0, 18, 6)], 1), "position": pos}
        reward = self.packed_reward_table()[((w1 >> 10) & 127).long()]
        terminated = ((w0 >> 24) & 1).bool()
        frame = ((w1 >> 17) & 32767) - 1
        info = self._make_info_dict(frame.to(torch.int32), torch.stack([bits(w0, 25, 3), bits(w0, 28, 3), bits(w1, 0, 5), bits(w1, 5, 5)], 1), obs)
        return obs, reward, terminated, torch.zeros_like(terminated), info

    # ------------------------------------------------------------------ state access, statistics
    def get_state(self, first: int = 0, count: Optional[int] = None) -> np.ndarray:
        """Expanded per-env state as a numpy structured array (fg_env_state)."""
        count = self.num_envs - first if count is None else count
        out = np.zeros(count, dtype=_capi.env_state_dtype())
        _capi.check(self._lib.fg_get_state(self._handle, int(first), int(count), C.c_void_p(out.ctypes.data)))
        return out

    def set_state(self, states: np.ndarray, first: int = 0):
        states = np.ascontiguousarray(states, dtype=_capi.env_state_dtype())
        _capi.check(self._lib.fg_set_state(self._handle, int(first), len(states), C.c_void_p(states.ctypes.data)))

    def save_battle_state(self, index: Optional[int] = None):
        """FootsiesEnv.save_battle_state (footsies.py:432-437, STATE_SAVE): the battle of env `index` as a
        FootsiesBattleState in the reference's schema; index=None returns the single state when num_envs == 1
        (the reference's call) and a list over all envs otherwise."""
        from .state import env_state_to_battle_state
        if index is None:
            states = [env_state_to_battle_state(r) for r in self.get_state()]
            return states[0] if self.num_envs == 1 else states
        return env_state_to_battle_state(self.get_state(int(index), 1)[0])

    def load_battle_state(self, battle_state, index: int = 0):
        """FootsiesEnv.load_battle_state (footsies.py:439-444, STATE_LOAD -> BattleCore.LoadState,
        BattleCore.cs:677-683): overwrite the fighters and the frame counter of env `index`; actors' held inputs,
        bot queues, RNG and the reward accumulator are untouched like in the game.  Accepts a FootsiesBattleState
        or its JSON string.  Like in the reference, the next dense reward compares the next state's guard bars with those
        of the last state received BEFORE the load and the episode's cumulative reward runs on across it (footsies.py:530,
        556-558): the rewards of a loaded battle are recomputed on the host until its episode ends.  With frame_skip > 1 or
        the fused frame skipping the kernel's own value is kept (it compares with the loaded guard bars)."""
        from .state import FootsiesBattleState, battle_state_into_env_state
        if isinstance(battle_state, str):
            battle_state = FootsiesBattleState.from_json(battle_state)
        index = int(index)
        rec = self.get_state(index, 1)
        guard_seen = rec[0]["f"]["guard"].copy()            # guard bars of the last state the agent received
        cum_index = int(rec[0]["cum_reward_index"])
        battle_state_into_env_state(battle_state, rec[0])
        self.set_state(rec, index)
        if self.dense_reward and self.frame_skip == 1 and not self.skip_unactionable:
            # footsies.py:530, 556-558: the Python side never learns about the load -- its next dense reward compares the
            # guard bars of the next state with those of the last state RECEIVED, and its cumulative episode reward runs on
            # (possibly past what one round can reach, which is all the kernel's 13-value reward automaton covers).  From
            # here to the end of this episode the battle's reward is therefore recomputed on the host in Python float
            # arithmetic, exactly like the reference does (_apply_load_fix); a handful of battles, never the hot path.
            from .reward_automaton import CUM_VALUES
            if index not in self._load_fix:
                self._load_fix[index] = {"cum": float(CUM_VALUES[cum_index]), "guard": (int(guard_seen[0]), int(guard_seen[1]))}

    def _apply_load_fix(self, host_reward=None):
        """Battles that went through load_battle_state: FootsiesEnv._get_dense_reward (footsies.py:388-405) in Python floats
        on the guard bars before / after this step, until the episode ends."""
        for index in list(self._load_fix):
            tr = self._load_fix[index]
            rec = self.get_state(index, 1)[0]
            if int(rec["frame"]) == -1:
                del self._load_fix[index]                  # the battle was restarted: the kernel's accumulator starts afresh
                continue
            g_now = (int(rec["f"]["guard"][0]), int(rec["f"]["guard"][1]))
            reward = 0.0
            if g_now[0] < tr["guard"][0]:
                reward -= 0.3
            if g_now[1] < tr["guard"][1]:
                reward += 0.3
            tr["cum"] += reward
            tr["guard"] = g_now
            if bool(rec["done"]):
                reward += (1 if int(rec["f"]["vital"][1]) == 0 else -1) - tr["cum"]
                del self._load_fix[index]
            r32 = float(np.float32(reward))
            self.reward[index] = r32
            if host_reward is not None:
                host_reward[index] = r32

    def episode_stats(self) -> dict:
        """Episode statistics accumulated on the device by the step kernel's warp reductions."""
        out = np.zeros(_capi.FG_STAT_COUNT, dtype=np.uint64)
        _capi.check(self._lib.fg_read_stats(self._handle, C.c_void_p(out.ctypes.data), self._stream()))
        return {k: int(v) for k, v in zip(_capi.STAT_NAMES, out)}

    def all_reduce_stats(self) -> dict:
        """End-of-rollout statistics summed over all ranks (the only collective on this path)."""
        from .distributed import all_reduce_stats
        return all_reduce_stats(self.stats)

    def launch_count(self) -> int:
        return int(self._lib.fg_launch_count(self._handle))

    # reference helper kept for completeness (footsies.py:590-614): there are no ports to find
    @staticmethod
    def find_ports(start: int, step: int = 1, stop=None):
        raise RuntimeError("find_ports is meaningless here: the simulator uses no sockets")


MOVE_DURATIONS = list(_fd.MOVE_DURATIONS)
