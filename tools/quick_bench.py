#!/usr/bin/env python3
"""Steady-state timing of the step kernel (developer tool): burn in until episodes are desynchronised, then time."""
import argparse, os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from footsies_gym_b200 import FootsiesEnv

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=4 * 1024 * 1024)
ap.add_argument("--burnin", type=int, default=600)
ap.add_argument("--steps", type=int, default=200)
ap.add_argument("--k", type=int, default=1)
ap.add_argument("--selfplay", action="store_true")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--graph", action="store_true", help="capture the timed steps into one CUDA graph (removes the Python launch overhead)")
a = ap.parse_args()
dev = torch.device("cuda", 0)
env = FootsiesEnv(num_envs=a.envs, device=dev, opponent="self_play" if a.selfplay else None, frame_skip=a.k, seed=0)
env.reset()
g = torch.Generator(device=dev); g.manual_seed(1234)
t1 = [torch.randint(0, 8, (a.envs,), generator=g, device=dev, dtype=torch.uint8) for _ in range(8)]
t2 = [torch.randint(0, 8, (a.envs,), generator=g, device=dev, dtype=torch.uint8) for _ in range(8)]
def step(i):
    env.bind_actions(t1[i % 8], t2[i % 8] if a.selfplay else None)
    env.step_bound()
for i in range(a.burnin): step(i)
torch.cuda.synchronize()
res = []
graph = None
if a.graph:
    st = torch.cuda.Stream(device=dev)
    st.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(st):
        for i in range(3): step(i)
    torch.cuda.current_stream(dev).wait_stream(st)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for i in range(a.steps): step(i)
for r in range(a.reps):
    f0 = env.episode_stats()["env_frames"]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if graph is not None:
        graph.replay()
    else:
        for i in range(a.steps): step(i)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1); fr = env.episode_stats()["env_frames"] - f0
    us = ms * 1e3 / a.steps
    gbs = env.algorithmic_bytes_per_env_step * a.envs / (us * 1e-6) / 1e9
    res.append(dict(us_per_step=round(us, 2), frames_per_s=round(fr / (ms * 1e-3) / 1e9, 3), alg_GBs=round(gbs, 1), frac=round(gbs / 6547.8, 4)))
print(json.dumps(dict(envs=a.envs, k=a.k, selfplay=a.selfplay, burnin=a.burnin, graph=a.graph, results=res)))
