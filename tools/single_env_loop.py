#!/usr/bin/env python3
"""BASELINE configs[0] on the drop-in: ONE FootsiesEnv used exactly like the reference's example loop (footsies.py:633-661: random
agent vs the in-game bot, tuple actions, reset() after termination), every observation read back to Python numbers each step.
The reference's game process is capped at 300 env-frames/s by its fixed timestep in fast-forward mode; this
measures what the Python + ctypes + launch + read-back overhead of a single-battle step costs here.
usage: python tools/single_env_loop.py [steps]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from footsies_gym_b200 import FootsiesEnv

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
env = FootsiesEnv(autoreset=False, seed=0)
rng = np.random.default_rng(0)
acts = rng.integers(0, 2, size=(steps, 3)).astype(bool)
obs, info = env.reset()
for t in range(200):
    obs, r, term, trunc, info = env.step(tuple(acts[t]))
    if bool(term[0]):
        env.reset()
torch.cuda.synchronize()
episodes, ret = 0, 0.0
t0 = time.perf_counter()
for t in range(steps):
    obs, r, term, trunc, info = env.step(tuple(acts[t]))
    x = obs["position"][0].tolist()            # an agent on the host looks at the observation ...
    g = obs["guard"][0].tolist()
    ret += float(r[0])
    if bool(term[0]):                           # ... and at the termination flag, every step
        episodes += 1
        obs, info = env.reset()
dt = time.perf_counter() - t0
print(f"single env, reference-style loop: {steps} steps in {dt:.2f} s = {steps / dt:.0f} env-steps/s ({dt / steps * 1e6:.1f} us per step), "
      f"{episodes} episodes, return sum {ret:+.1f}; the reference's game process: <= 300 env-frames/s")
env.close()
