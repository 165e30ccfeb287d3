#!/usr/bin/env python3
"""FootsiesFrameSkipped: fused into the step kernel (fg_config.skip_unactionable) vs the wrapper's loop of masked steps
(developer timing tool)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from footsies_gym_b200 import FootsiesEnv
from footsies_gym_b200.wrappers import FootsiesFrameSkipped

for n in (4096, 65536, 1048576):
    for fused in (True, False):
        env = FootsiesFrameSkipped(FootsiesEnv(num_envs=n, seed=0), fused=fused)
        env.reset()
        a = [torch.randint(0, 8, (n,), device="cuda", dtype=torch.uint8) for _ in range(4)]
        for i in range(10):
            env.step(a[i % 4])
        torch.cuda.synchronize()
        f0 = env.unwrapped.episode_stats()["env_frames"]
        t0 = time.perf_counter()
        steps = 100
        for i in range(steps):
            env.step(a[i % 4])
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        fr = env.unwrapped.episode_stats()["env_frames"] - f0
        print(f"n={n} fused={fused}: {dt / steps * 1e6:.0f} us per wrapper step, {fr / steps / n:.2f} frames per step, "
              f"{fr / dt:.3e} env-frames/s", flush=True)
        env.close()
