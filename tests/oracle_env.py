"""CPU stand-in with the FootsiesEnv surface the wrappers use, driven by the CPU oracle.

TEST INFRASTRUCTURE ONLY: lets the batched wrappers (pure torch ops) be checked on a machine without a GPU.
The product never imports this."""
import torch

import oracle_binding as ob
from footsies_gym_b200.env import FootsiesEnv, _as_bitmask
from footsies_gym_b200.moves import FootsiesMove
from footsies_gym_b200.spaces import footsies_action_space, footsies_observation_space


class OracleTorchEnv:
    is_base_footsies_env = True

    def __init__(self, num_envs=1, dense_reward=True, autoreset=False, seed=0, p2_bot=True):
        self.num_envs = num_envs
        self.device = torch.device("cpu")
        self.b = ob.OracleBatch(num_envs, p2_bot=p2_bot, dense_reward=dense_reward, autoreset=autoreset, seed=seed)
        relevant = [m for m in FootsiesMove if m.name not in ("WIN", "DEAD")]
        self.observation_space = footsies_observation_space(len(relevant), max(m.value.duration for m in relevant))
        self.action_space = footsies_action_space()
        self.obs = torch.zeros((num_envs, 8), dtype=torch.float32)
        self.reward = torch.zeros(num_envs, dtype=torch.float32)
        self.terminated = torch.zeros(num_envs, dtype=torch.bool)
        self.truncated = torch.zeros(num_envs, dtype=torch.bool)
        self.info_frame = torch.zeros(num_envs, dtype=torch.int32)
        self._obs_dict = FootsiesEnv._make_obs_dict(self.obs)
        self._mask = None

    def _publish(self, sel):
        t = self.b.trace
        self.obs[sel] = torch.from_numpy(t["obs"].copy())[sel]
        self.reward[sel] = torch.from_numpy(t["reward"].copy())[sel]
        self.terminated[sel] = torch.from_numpy(t["terminated"].astype(bool))[sel]
        self.info_frame[sel] = torch.from_numpy(t["info_frame"].copy())[sel]
        return self._obs_dict, {"frame": self.info_frame, **self._obs_dict}

    def reset(self, *, seed=None, options=None):
        if seed is not None:
            self.b.seed(seed)
        self.b.reset()
        return self._publish(slice(None))

    def set_step_mask(self, mask):
        self._mask = None if mask is None else mask.bool().clone()

    def step(self, action):
        a = _as_bitmask(action, self.num_envs, "cpu").numpy()
        if self._mask is None:
            self.b.step(a)
            sel = slice(None)
        else:
            # only masked envs advance: step a scratch copy env by env (N is tiny in these tests)
            assert self.num_envs == 1, "masked stepping of the CPU stand-in supports one env"
            if bool(self._mask[0]):
                self.b.step(a)
            sel = self._mask
        obs, info = self._publish(sel)
        return obs, self.reward, self.terminated, self.truncated, info

    def close(self):
        pass
