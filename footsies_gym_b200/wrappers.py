"""Batched tensor versions of the reference's Gymnasium wrappers (footsies-gym/footsies_gym/wrappers/*.py).

Same class names and semantics, applied to all N envs at once with torch ops on the env's device:
  FootsiesNormalized                       wrappers/normalization.py:6-55
  FootsiesActionCombinationsDiscretized    wrappers/action_comb_disc.py:5-18
  FootsiesFrameSkipped                     wrappers/frame_skip.py:6-80
  FootsiesStatistics                       wrappers/statistics.py:5-70
They wrap `footsies_gym_b200.FootsiesEnv` (or each other in the order the reference prescribes) and do not
depend on gymnasium.
"""
from typing import Optional

import torch

from .env import FootsiesEnv
from .moves import FOOTSIES_MOVE_INDEX_TO_MOVE, FootsiesMove
from .spaces import Box, Dict, Discrete


class _Wrapper:
    def __init__(self, env):
        self.env = env

    def __getattr__(self, name):          # forward everything else (num_envs, device, obs, episode_stats, ...)
        return getattr(self.env, name)

    @property
    def unwrapped(self):
        e = self.env
        while isinstance(e, _Wrapper):
            e = e.env
        return e

    def reset(self, *, seed=None, options=None):
        return self.env.reset(seed=seed, options=options)

    def step(self, action, *args, **kw):
        return self.env.step(action, *args, **kw)

    def close(self):
        return self.env.close()


class FootsiesNormalized(_Wrapper):
    """All observation variables in [0, 1] / [-1, 1]: guard / 3, position / 4.6, move_frame / move duration.
    Must wrap the base environment, before any other observation wrapper (normalization.py:15-19)."""

    def __init__(self, env, normalize_guard: bool = True):
        if not (isinstance(env, FootsiesEnv) or getattr(env, "is_base_footsies_env", False)):
            raise ValueError("FootsiesNormalized wrapper should be applied to the base FOOTSIES environment")
        super().__init__(env)
        self.normalize_guard = normalize_guard
        sp = dict(env.observation_space.spaces)
        if normalize_guard:
            sp["guard"] = Box(low=0.0, high=1.0, shape=(2,))
        sp["move_frame"] = Box(low=0.0, high=1.0, shape=(2,))
        sp["position"] = Box(low=-1.0, high=1.0, shape=(2,))
        self.observation_space = Dict(sp)
        self._durations = torch.tensor([float(m.value.duration) for m in FOOTSIES_MOVE_INDEX_TO_MOVE],
                                       dtype=torch.float32, device=env.device)
        self._out = torch.zeros((env.num_envs, 8), dtype=torch.float32, device=env.device)
        self._out_dict = FootsiesEnv._make_obs_dict(self._out)

    def observation(self, obs: dict) -> dict:
        o = self._out
        if self.normalize_guard:
            torch.div(obs["guard"], 3.0, out=o[:, 0:2])
        else:
            o[:, 0:2].copy_(obs["guard"])
        o[:, 2:4].copy_(obs["move"])
        torch.div(obs["move_frame"], self._durations[obs["move"].long()], out=o[:, 4:6])
        torch.div(obs["position"], 4.6, out=o[:, 6:8])
        return self._out_dict

    def reset(self, *, seed=None, options=None):
        obs, info = self.env.reset(seed=seed, options=options)
        return self.observation(obs), info

    def step(self, action, *args, **kw):
        obs, reward, terminated, truncated, info = self.env.step(action, *args, **kw)
        return self.observation(obs), reward, terminated, truncated, info

    @staticmethod
    def undo(obs: dict, normalized_guard: bool = True) -> dict:
        obs = dict(obs)
        dur = torch.tensor([float(m.value.duration) for m in FOOTSIES_MOVE_INDEX_TO_MOVE], dtype=torch.float32,
                           device=obs["move"].device)
        if normalized_guard:
            obs["guard"] = obs["guard"] * 3.0
        obs["position"] = obs["position"] * 4.6
        obs["move_frame"] = obs["move_frame"] * dur[obs["move"].long()]
        return obs


class FootsiesActionCombinationsDiscretized(_Wrapper):
    """Discrete(8) actions: bit 0 left, bit 1 right, bit 2 attack -- the game's own input encoding, which is also
    the kernel's native action format, so this wrapper only declares the space and casts."""

    def __init__(self, env):
        super().__init__(env)
        self.action_space = Discrete(2 ** 3)

    def action(self, act):
        if not isinstance(act, torch.Tensor):
            act = torch.as_tensor(act)
        return act.reshape(-1).to(torch.uint8)

    def step(self, action, *args, **kw):
        return self.env.step(self.action(action), *args, **kw)


_HIT_GUARD_MOVES = (FootsiesMove.DAMAGE, FootsiesMove.GUARD_STAND, FootsiesMove.GUARD_CROUCH, FootsiesMove.GUARD_M,
                    FootsiesMove.GUARD_BREAK)


class FootsiesFrameSkipped(_Wrapper):
    """Skip the time steps on which the agent cannot act (frame_skip.py:46-80): envs whose observation is
    skippable keep stepping with the no-op action and the rewards are accumulated.  P1's move_frame is dropped from the
    observation.  With the CUDA env against the in-game bot the whole loop runs inside the step kernel
    (fg_config.skip_unactionable: one launch per step, no host round trip); otherwise it is a loop of masked steps here --
    only the skippable envs advance (the kernel's step mask; the frame_delay queue of a held-back env stands still).
    When P2's actions come from outside, the action given to step() is held over the skipped frames (see __init__)."""

    def __init__(self, env, fused: Optional[bool] = None):
        super().__init__(env)
        base0 = self.unwrapped
        can_fuse = hasattr(base0, "set_skip_unactionable") and getattr(base0, "opponent", None) is None \
            and not getattr(base0, "by_example", False) and getattr(base0, "frame_delay", 0) == 0
        if fused and not can_fuse:
            raise ValueError("fused frame skipping needs the CUDA env, an agent-controlled P1, no frame_delay and an "
                             "opponent that is not a Python callable")
        # Default: fuse only against the in-game bot.  With P2 driven from outside (opponent="self_play" / "remote") the
        # reference's wrapper asks the opponent again on every skipped step (frame_skip.py:72 -> footsies.py:522-527), which
        # a single launch cannot do: there both paths HOLD the P2 action given to step() for the skipped frames (the masked
        # loop re-passes it), so fusing is only done on request (fused=True).
        bot_p2 = getattr(base0, "_opponent_mode", "bot") == "bot"
        self.fused = (can_fuse and bot_p2) if fused is None else bool(fused)
        if self.fused:
            base0.set_skip_unactionable(True)
        sp = dict(env.observation_space.spaces)
        mf = sp["move_frame"]
        sp["move_frame"] = Box(low=float(mf.low[1]), high=float(mf.high[1]), shape=(1,))
        self.observation_space = Dict(sp)
        base = self.unwrapped
        idx = {m: i for i, m in enumerate(FOOTSIES_MOVE_INDEX_TO_MOVE)}
        hg = torch.zeros(len(idx), dtype=torch.bool, device=base.device)
        for m in _HIT_GUARD_MOVES:
            hg[idx[m]] = True
        self._hit_guard = hg
        self._damage_idx = idx[FootsiesMove.DAMAGE]
        self._retained = torch.zeros(base.num_envs, dtype=torch.float32, device=base.device)
        self._zero_action = torch.zeros(base.num_envs, dtype=torch.uint8, device=base.device)

    @staticmethod
    def _frame_skip_obs(obs: dict) -> dict:
        return {"guard": obs["guard"], "move": obs["move"], "move_frame": obs["move_frame"][:, 1:2],
                "position": obs["position"]}

    def _is_obs_skippable(self, obs: dict) -> torch.Tensor:
        move = obs["move"].long()
        return ((obs["move_frame"][:, 0] != 0.0) & ~self._hit_guard[move[:, 1]]) | (move[:, 0] == self._damage_idx)

    def reset(self, *, seed=None, options=None):
        obs, info = self.env.reset(seed=seed, options=options)
        return self._frame_skip_obs(obs), info

    def step(self, action, *args, **kw):
        base = self.unwrapped
        obs, reward, terminated, truncated, info = self.env.step(action, *args, **kw)
        if self.fused:
            return self._frame_skip_obs(obs), reward, terminated, truncated, info
        self._retained.copy_(reward)
        skip = self._is_obs_skippable(obs) & ~(terminated | truncated)
        while bool(skip.any()):
            base.set_step_mask(skip)
            try:
                obs, reward, terminated, truncated, info = self.env.step(self._zero_action, *args, **kw)
            finally:
                base.set_step_mask(None)
            self._retained += torch.where(skip, reward, torch.zeros_like(reward))
            skip = skip & self._is_obs_skippable(obs) & ~(terminated | truncated)
        return self._frame_skip_obs(obs), self._retained, terminated, truncated, info


class FootsiesStatistics(_Wrapper):
    """Special-move counters per episode (statistics.py:26-50), kept per env on the device; finished episodes are
    appended to the metric lists.  (The kernel also accumulates the totals: FootsiesEnv.episode_stats().)"""

    def __init__(self, env):
        super().__init__(env)
        base = self.unwrapped
        idx = {m: i for i, m in enumerate(FOOTSIES_MOVE_INDEX_TO_MOVE)}
        self._special = torch.zeros(len(idx), dtype=torch.bool, device=base.device)
        self._normal = torch.zeros(len(idx), dtype=torch.bool, device=base.device)
        self._special[idx[FootsiesMove.B_SPECIAL]] = self._special[idx[FootsiesMove.N_SPECIAL]] = True
        self._normal[idx[FootsiesMove.B_ATTACK]] = self._normal[idx[FootsiesMove.N_ATTACK]] = True
        n = base.num_envs
        self._counter = torch.zeros(n, dtype=torch.int64, device=base.device)
        self._neutral_counter = torch.zeros(n, dtype=torch.int64, device=base.device)
        self._prev_p1_move: Optional[torch.Tensor] = None
        self._special_moves_per_episode = []
        self._special_moves_from_neutral_per_episode = []

    def reset(self, *, seed=None, options=None):
        obs, info = self.env.reset(seed=seed, options=options)
        self._prev_p1_move = obs["move"][:, 0].long().clone()
        return obs, info

    def step(self, action, *args, **kw):
        obs, reward, terminated, truncated, info = self.env.step(action, *args, **kw)
        p1 = obs["move"][:, 0].long()
        new_special = (self._prev_p1_move != p1) & self._special[p1]
        self._counter += new_special
        self._neutral_counter += new_special & ~self._normal[self._prev_p1_move]
        self._prev_p1_move = p1.clone()
        done = terminated | truncated
        if bool(done.any()):
            self._special_moves_per_episode += self._counter[done].tolist()
            # the reference never appends to the from-neutral list (statistics.py:44-48 only bumps the counter);
            # kept as is so reports match
            self._counter[done] = 0
        return obs, reward, terminated, truncated, info

    @property
    def metric_special_moves_per_episode(self):
        return self._special_moves_per_episode

    @property
    def metric_special_moves_from_neutral_per_episode(self):
        return self._special_moves_from_neutral_per_episode
