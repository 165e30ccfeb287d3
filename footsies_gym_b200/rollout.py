"""On-device rollout collection (BASELINE.json configs[4]: a torch MLP policy consuming `env.obs` in place).

The reference's training loops step one socket-connected game per process and move every observation through JSON
(footsies.py:518-570).  Here a whole horizon of policy-forward -> sample -> FootsiesEnv.step for N battles is a
sequence of device kernels with no host round trip: the policy reads the kernel's observation tensor directly, writes
its sampled actions into the uint8 tensor the step kernel is bound to, and the per-step transition is appended to
pre-allocated [horizon, N, ...] buffers.  With `use_cuda_graph=True` the horizon is captured once into a CUDA graph
and replayed, which removes the per-step launch latency that dominates at 16k envs per GPU.

PyTorch is plumbing here (the policy is the user's model); the simulator step inside the loop is fg_step.
"""
from typing import Callable, Optional

import torch

from .env import FootsiesEnv

# FootsiesNormalized scales (wrappers/normalization.py:28-55): guard / 3, move index as is, move_frame / 55, position / 4.6
_OBS_SCALE = (1 / 3.0, 1 / 3.0, 1 / 14.0, 1 / 14.0, 1 / 55.0, 1 / 55.0, 1 / 4.6, 1 / 4.6)


class MLPPolicy(torch.nn.Module):
    """8 -> hidden -> hidden -> 8 logits over the 2^3 input combinations (wrappers/action_comb_disc.py:13-18)."""

    def __init__(self, hidden: int = 64):
        super().__init__()
        self.register_buffer("scale", torch.tensor(_OBS_SCALE, dtype=torch.float32))
        self.net = torch.nn.Sequential(torch.nn.Linear(8, hidden), torch.nn.Tanh(), torch.nn.Linear(hidden, hidden),
                                       torch.nn.Tanh(), torch.nn.Linear(hidden, 8))

    def forward(self, obs: torch.Tensor) -> torch.Tensor:
        return self.net(obs * self.scale)


class RolloutCollector:
    """Collects `horizon` steps of (obs, action, log-prob, reward, done) for every env of one GPU, entirely on device."""

    def __init__(self, env: FootsiesEnv, policy: Callable[[torch.Tensor], torch.Tensor], horizon: int = 128,
                 use_cuda_graph: bool = True, generator: Optional[torch.Generator] = None):
        if env.by_example:
            raise ValueError("the policy drives P1: create the env with by_example=False")
        self.env, self.policy, self.horizon = env, policy, int(horizon)
        n, dev = env.num_envs, env.device
        self.obs = torch.zeros((horizon, n, 8), dtype=torch.float32, device=dev)
        self.actions = torch.zeros((horizon, n), dtype=torch.uint8, device=dev)
        self.logp = torch.zeros((horizon, n), dtype=torch.float32, device=dev)
        self.rewards = torch.zeros((horizon, n), dtype=torch.float32, device=dev)
        self.dones = torch.zeros((horizon, n), dtype=torch.bool, device=dev)
        self._act = torch.zeros(n, dtype=torch.uint8, device=dev)     # the tensor the step kernel reads its actions from
        env.bind_actions(self._act)
        self._generator = generator
        self._graph = None
        self.use_cuda_graph = bool(use_cuda_graph)
        if not env.has_reset:
            env.reset()

    @torch.no_grad()
    def _one_horizon(self):
        env = self.env
        for t in range(self.horizon):
            self.obs[t].copy_(env.obs)                                 # the kernel overwrites env.obs in place next step
            logits = self.policy(env.obs)
            logp_all = torch.log_softmax(logits, dim=-1)
            a = torch.multinomial(logp_all.exp(), 1, generator=self._generator).squeeze(1)
            self.logp[t].copy_(logp_all.gather(1, a.unsqueeze(1)).squeeze(1))
            self._act.copy_(a)                                         # int64 -> uint8 bitmask (Left 1 | Right 2 | Attack 4)
            self.actions[t].copy_(self._act)
            env.step_bound()
            self.rewards[t].copy_(env.reward)
            self.dones[t].copy_(env.terminated)

    def collect(self):
        """Runs one horizon; returns the rollout buffers (views, overwritten by the next call)."""
        if not self.use_cuda_graph:
            self._one_horizon()
        else:
            if self._graph is None:
                stream = torch.cuda.Stream(device=self.env.device)
                stream.wait_stream(torch.cuda.current_stream(self.env.device))
                with torch.cuda.stream(stream):
                    self._one_horizon()                                # warm-up outside the capture (lazy inits)
                torch.cuda.current_stream(self.env.device).wait_stream(stream)
                torch.cuda.synchronize(self.env.device)
                self._graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self._graph):
                    self._one_horizon()
            self._graph.replay()
        return {"obs": self.obs, "actions": self.actions, "logp": self.logp, "rewards": self.rewards,
                "dones": self.dones, "last_obs": self.env.obs}
