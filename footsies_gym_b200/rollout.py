"""On-device rollout collection (BASELINE.json configs[4]: an MLP policy consuming the observation tensor in place).

The reference's training loops step one socket-connected game per process and move every observation through JSON
(footsies.py:518-570, 633-661).  Here a whole horizon of policy-forward -> sample -> FootsiesEnv.step for N battles is
a sequence of device kernels with no host round trip, captured once into a CUDA graph and replayed.

Three policy paths:
  * any torch callable `policy(obs[N, 8]) -> logits[N, 8]` (a dozen small torch kernels per step);
  * `MLPPolicy`, fused="step": one hand-written kernel per step (csrc/policy_kernel.cu, fg_policy_mlp_sample) does
    scale -> 8-H-H-8 tanh MLP -> log-softmax -> sample, and the rollout needs no copies at all: the step kernel is
    re-bound, per captured launch, to write observation t + 1, reward t and done t straight into the rollout buffers,
    and the policy kernel reads observation t from there.  Per step: 2 launches (policy, simulator step).
  * `MLPPolicy`, fused="horizon" (the default whenever the env qualifies: P1 = policy, P2 = in-game bot or a second
    MLPPolicy given as `opponent_policy` (self-play), autoreset, no step mask): ONE launch per horizon (csrc/rollout_kernel.cu, fg_rollout_mlp) -- a CTA keeps its 64 battles in
    registers from the first step to the last and alternates policy inference and the frame update; bit-identical
    to the per-step path.

PyTorch is plumbing here (parameters, buffers, CUDA graph); the simulator step inside the loop is fg_step.
"""
import ctypes as C
from typing import Callable, Optional

import torch

from . import _capi
from .env import FootsiesEnv

# FootsiesNormalized scales (wrappers/normalization.py:28-55): guard / 3, move index / 14, move_frame / 55, position / 4.6
_OBS_SCALE = (1 / 3.0, 1 / 3.0, 1 / 14.0, 1 / 14.0, 1 / 55.0, 1 / 55.0, 1 / 4.6, 1 / 4.6)


class MLPPolicy(torch.nn.Module):
    """8 -> hidden -> hidden -> 8 logits over the 2^3 input combinations (wrappers/action_comb_disc.py:13-18).
    hidden in {32, 64, 128} can be run by the fused inference kernel."""

    def __init__(self, hidden: int = 64):
        super().__init__()
        self.hidden = int(hidden)
        self.register_buffer("scale", torch.tensor(_OBS_SCALE, dtype=torch.float32))
        self.net = torch.nn.Sequential(torch.nn.Linear(8, hidden), torch.nn.Tanh(), torch.nn.Linear(hidden, hidden),
                                       torch.nn.Tanh(), torch.nn.Linear(hidden, 8))

    def forward(self, obs: torch.Tensor) -> torch.Tensor:
        return self.net(obs * self.scale)

    def fused_sample(self, obs: torch.Tensor, actions: torch.Tensor, logp: Optional[torch.Tensor], seed: int, counter: int,
                     counter_base: Optional[torch.Tensor] = None, obs_copy: Optional[torch.Tensor] = None,
                     first_env_index: int = 0):
        """One launch of fg_policy_mlp_sample on the current stream: actions (uint8 [N]) and logp (float32 [N]) are
        written in place; sampling is a pure function of (seed, counter + counter_base[0], first_env_index + env index)
        -- the global battle index, so that shards of one job draw what the whole batch would; counter_base is an optional
        int64 device tensor (bumped between CUDA-graph replays)."""
        lib = _capi.load()
        l1, l2, l3 = self.net[0], self.net[2], self.net[4]
        ts = [obs, self.scale, l1.weight, l1.bias, l2.weight, l2.bias, l3.weight, l3.bias]
        for t in ts:
            if t.dtype != torch.float32 or not t.is_contiguous() or t.device != obs.device:
                raise ValueError("fused policy inference needs contiguous float32 tensors on one device")
        if actions.dtype != torch.uint8 or not actions.is_contiguous():
            raise ValueError("actions must be a contiguous uint8 tensor")
        ptr = lambda t: None if t is None else C.c_void_p(t.data_ptr())   # noqa: E731
        rc = lib.fg_policy_mlp_sample(*[ptr(t) for t in ts], self.hidden, obs.shape[0], int(seed) & (2**64 - 1),
                                      int(counter), ptr(counter_base), ptr(actions), ptr(logp), ptr(obs_copy),
                                      int(first_env_index), C.c_void_p(torch.cuda.current_stream(obs.device).cuda_stream))
        if rc != 0:
            raise _capi.FootsiesLibraryError(lib.fg_policy_last_error().decode())

    def fused_sample_p2(self, obs: torch.Tensor, actions: torch.Tensor, logp: Optional[torch.Tensor], seed: int, counter: int,
                        counter_base: Optional[torch.Tensor] = None, mirror: bool = False, first_env_index: int = 0):
        """fused_sample for a policy that drives P2 (fg_policy_mlp_sample_p2): with mirror=True it is fed the mirrored
        observation (per-player fields swapped, positions negated) and its Left / Right bits are mirrored back, so that
        the network that plays P1 can play P2 as well."""
        lib = _capi.load()
        l1, l2, l3 = self.net[0], self.net[2], self.net[4]
        ts = [obs, self.scale, l1.weight, l1.bias, l2.weight, l2.bias, l3.weight, l3.bias]
        for t in ts:
            if t.dtype != torch.float32 or not t.is_contiguous() or t.device != obs.device:
                raise ValueError("fused policy inference needs contiguous float32 tensors on one device")
        if actions.dtype != torch.uint8 or not actions.is_contiguous():
            raise ValueError("actions must be a contiguous uint8 tensor")
        ptr = lambda t: None if t is None else C.c_void_p(t.data_ptr())   # noqa: E731
        rc = lib.fg_policy_mlp_sample_p2(*[ptr(t) for t in ts], self.hidden, obs.shape[0], int(seed) & (2**64 - 1),
                                         int(counter), ptr(counter_base), ptr(actions), ptr(logp), int(bool(mirror)),
                                         int(first_env_index), C.c_void_p(torch.cuda.current_stream(obs.device).cuda_stream))
        if rc != 0:
            raise _capi.FootsiesLibraryError(lib.fg_policy_last_error().decode())


def mirror_obs(obs: torch.Tensor) -> torch.Tensor:
    """The observation [N, 8] as P2 sees it: per-player fields swapped, positions negated."""
    m = obs[:, [1, 0, 3, 2, 5, 4, 7, 6]].clone()
    m[:, 6:8] = -m[:, 6:8]
    return m


_MIRROR_ACTION = (0, 2, 1, 3, 4, 6, 5, 7)      # Left (1) <-> Right (2)
_P2_SEED_SALT = 0x5DEECE66D2F1A3B7


class RolloutCollector:
    """Collects `horizon` steps of (obs, action, log-prob, reward, done) for every env of one GPU, entirely on device.

    Buffers (overwritten by every collect()): obs [horizon + 1, N, 8] (obs[t] is what the policy saw at step t, obs[-1]
    the observation after the last step), actions uint8 [horizon, N], logp / rewards float32 [horizon, N], dones bool."""

    def __init__(self, env: FootsiesEnv, policy: Callable[[torch.Tensor], torch.Tensor], horizon: int = 128,
                 use_cuda_graph: bool = True, fused=None, seed: int = 0,
                 opponent_policy: Optional[Callable[[torch.Tensor], torch.Tensor]] = None, mirror_opponent: bool = False):
        """fused: None / True = the most fused path available ("horizon", else "step", else torch ops for a policy that
        is not an MLPPolicy); False = torch ops; "step" / "horizon" = that path or an error.
        opponent_policy: a second policy for P2 (self-play rollouts; the env must take P2's actions from step(), i.e.
        opponent="self_play" / "remote").  It sees the same observation as the reference's `opponent(obs, info)`
        callable (footsies.py:522-527), or its mirror image with mirror_opponent=True (then `policy` itself can be passed
        again: one network plays both sides).  Its actions / log-probabilities are collected in actions_p2 / logp_p2."""
        if env.by_example:
            raise ValueError("the policy drives P1: create the env with by_example=False")
        if env.frame_delay:
            raise ValueError("RolloutCollector binds the env's outputs directly: frame_delay must be 0")
        self.env, self.policy, self.horizon = env, policy, int(horizon)
        self.opponent_policy, self.mirror_opponent = opponent_policy, bool(mirror_opponent)
        if opponent_policy is not None and (env._opponent_mode == "bot" or env.opponent is not None):
            raise ValueError("opponent_policy needs an env whose P2 actions come from step(): opponent='self_play'")
        can_step = isinstance(policy, MLPPolicy) and policy.hidden in (32, 64, 128)
        if opponent_policy is not None:
            can_step = can_step and isinstance(opponent_policy, MLPPolicy) and opponent_policy.hidden == policy.hidden
        can_horizon = (can_step and env.autoreset and env._step_mask is None
                       and (env._opponent_mode == "bot" or opponent_policy is not None))
        if fused is None or fused is True:
            if fused and not can_step:
                raise ValueError("fused=True needs an MLPPolicy with hidden in {32, 64, 128}")
            self.mode = "horizon" if can_horizon else "step" if can_step else "torch"
        elif fused is False:
            self.mode = "torch"
        elif fused in ("step", "horizon"):
            if not (can_horizon if fused == "horizon" else can_step):
                raise ValueError(f"fused={fused!r} is not available for this policy / env configuration")
            self.mode = fused
        else:
            raise ValueError("fused must be None, a bool, 'step' or 'horizon'")
        self.fused = self.mode != "torch"
        n, dev, h = env.num_envs, env.device, self.horizon
        self.obs = torch.zeros((h + 1, n, 8), dtype=torch.float32, device=dev)
        self.actions = torch.zeros((h, n), dtype=torch.uint8, device=dev)
        self.logp = torch.zeros((h, n), dtype=torch.float32, device=dev)
        self.rewards = torch.zeros((h, n), dtype=torch.float32, device=dev)
        self.dones = torch.zeros((h, n), dtype=torch.bool, device=dev)
        self.actions_p2 = self.logp_p2 = None
        if opponent_policy is not None:
            self.actions_p2 = torch.zeros((h, n), dtype=torch.uint8, device=dev)
            self.logp_p2 = torch.zeros((h, n), dtype=torch.float32, device=dev)
            self._mirror_action = torch.tensor(_MIRROR_ACTION, dtype=torch.uint8, device=dev)
        self._seed = int(seed)
        self._drawn = torch.zeros(1, dtype=torch.int64, device=dev)   # policy steps taken so far (device side: graph replays bump it)
        self._graph = None
        self.use_cuda_graph = bool(use_cuda_graph)
        if not env.has_reset:
            env.reset()
        self.obs[h].copy_(env.obs)                                  # the observation the first step will act on

    def _horizon_launch(self):
        """fg_rollout_mlp: the whole horizon in one launch on the current stream."""
        env, pol = self.env, self.policy
        l1, l2, l3 = pol.net[0], pol.net[2], pol.net[4]
        r = _capi.FgRolloutBuffers(struct_size=C.sizeof(_capi.FgRolloutBuffers), hidden=pol.hidden, horizon=self.horizon,
                                   reserved0=0, seed=self._seed & (2**64 - 1))
        ts = dict(scale=pol.scale, w1=l1.weight, b1=l1.bias, w2=l2.weight, b2=l2.bias, w3=l3.weight, b3=l3.bias,
                  counter_base=self._drawn, obs=self.obs, actions=self.actions, logp=self.logp, rewards=self.rewards,
                  dones=self.dones)
        if self.opponent_policy is not None:
            o = self.opponent_policy
            m1, m2, m3 = o.net[0], o.net[2], o.net[4]
            ts.update(p2_scale=o.scale, p2_w1=m1.weight, p2_b1=m1.bias, p2_w2=m2.weight, p2_b2=m2.bias, p2_w3=m3.weight,
                      p2_b3=m3.bias, actions_p2=self.actions_p2, logp_p2=self.logp_p2)
            r.p2_seed = (self._seed ^ _P2_SEED_SALT) & (2**64 - 1)
            r.p2_mirror = int(self.mirror_opponent)
        for k, t in ts.items():
            if not t.is_contiguous() or t.device != env.device:
                raise ValueError(f"fused rollout needs contiguous tensors on the env's device ({k})")
            setattr(r, k, t.data_ptr())
        _capi.check(env._lib.fg_rollout_mlp(env._handle, C.byref(r), env._stream()))
        self._drawn.add_(self.horizon)

    @torch.no_grad()
    def _one_horizon(self):
        env, h = self.env, self.horizon
        if self.mode == "horizon":
            return self._horizon_launch()
        self.obs[0].copy_(self.obs[h])                              # carry the last observation over
        for t in range(h):
            if self.mode == "step":
                self.policy.fused_sample(self.obs[t], self.actions[t], self.logp[t], self._seed, t, self._drawn,
                                         first_env_index=env.first_env_index)
            else:
                logits = self.policy(self.obs[t])
                logp_all = torch.log_softmax(logits, dim=-1)
                a = torch.multinomial(logp_all.exp(), 1).squeeze(1)
                self.logp[t].copy_(logp_all.gather(1, a.unsqueeze(1)).squeeze(1))
                self.actions[t].copy_(a)                            # int64 -> uint8 bitmask (Left 1 | Right 2 | Attack 4)
            if self.opponent_policy is None:
                env.bind_actions(self.actions[t])
            else:
                if self.mode == "step":
                    self.opponent_policy.fused_sample_p2(self.obs[t], self.actions_p2[t], self.logp_p2[t],
                                                         self._seed ^ _P2_SEED_SALT, t, self._drawn, self.mirror_opponent,
                                                         first_env_index=env.first_env_index)
                else:
                    o = mirror_obs(self.obs[t]) if self.mirror_opponent else self.obs[t]
                    logp_all = torch.log_softmax(self.opponent_policy(o), dim=-1)
                    a = torch.multinomial(logp_all.exp(), 1).squeeze(1)
                    self.logp_p2[t].copy_(logp_all.gather(1, a.unsqueeze(1)).squeeze(1))
                    self.actions_p2[t].copy_(self._mirror_action[a] if self.mirror_opponent else a)
                env.bind_actions(self.actions[t], self.actions_p2[t])
            env.bind_outputs(obs=self.obs[t + 1], reward=self.rewards[t], terminated=self.dones[t])
            env.step_bound()
        self._drawn.add_(h)

    def collect(self):
        """Runs one horizon; returns the rollout buffers (views, overwritten by the next call)."""
        dev = self.env.device
        if not self.use_cuda_graph or self.mode == "horizon":      # one launch: nothing for a graph to save
            self._one_horizon()
        else:
            if self._graph is None:
                stream = torch.cuda.Stream(device=dev)
                stream.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(stream):
                    self._one_horizon()                             # warm-up outside the capture (lazy inits)
                torch.cuda.current_stream(dev).wait_stream(stream)
                torch.cuda.synchronize(dev)
                self._graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self._graph):
                    self._one_horizon()
            self._graph.replay()
        out = {"obs": self.obs[:self.horizon], "actions": self.actions, "logp": self.logp, "rewards": self.rewards,
               "dones": self.dones, "last_obs": self.obs[self.horizon]}
        if self.opponent_policy is not None:
            out.update(actions_p2=self.actions_p2, logp_p2=self.logp_p2)
        return out
