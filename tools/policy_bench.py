#!/usr/bin/env python3
"""Timing of the fused policy kernel alone and of a rollout step (developer tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from footsies_gym_b200 import FootsiesEnv
from footsies_gym_b200.rollout import MLPPolicy, RolloutCollector
dev = torch.device("cuda:0")
for n in (16384, 262144):
    for hidden in (64, 128):
        pol = MLPPolicy(hidden).to(dev)
        obs = torch.rand(n, 8, device=dev)
        act = torch.zeros(n, dtype=torch.uint8, device=dev); lp = torch.zeros(n, device=dev)
        for _ in range(5): pol.fused_sample(obs, act, lp, 1, 0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(200): pol.fused_sample(obs, act, lp, 1, i)
        e1.record(); torch.cuda.synchronize()
        print(f"policy kernel n={n} hidden={hidden}: {e0.elapsed_time(e1) * 1e3 / 200:.2f} us")
for n, hidden, mode in ((16384, 64, "step"), (16384, 64, "horizon"), (16384, 32, "horizon"), (16384, 128, "horizon"),
                        (131072, 64, "horizon"), (1048576, 64, "horizon")):
    env = FootsiesEnv(num_envs=n, device=dev, seed=0)
    pol = MLPPolicy(hidden).to(dev)
    col = RolloutCollector(env, pol, horizon=128, use_cuda_graph=True, fused=mode)
    col.collect(); col.collect(); torch.cuda.synchronize()
    f0 = env.episode_stats()["env_frames"]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): col.collect()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    fr = env.episode_stats()["env_frames"] - f0
    print(f"rollout {n} x 128 hidden={hidden} mode={mode}: {ms:.3f} ms per horizon = {ms * 1e3 / 128:.2f} us per step, "
          f"{fr / 5 / (ms * 1e-3):.3e} env-frames/s")
    env.close(); del col, env
