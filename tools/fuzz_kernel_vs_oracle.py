#!/usr/bin/env python3
"""Differential fuzzing of the frame logic of the CUDA path against the CPU oracle beyond the fixed tapes of the
test-suite: randomly drawn configurations (who is a bot, reward mode, fused frame-skip K, autoreset, stale Intro input,
seeds, global index offsets, ragged batch sizes), input personalities and masked mid-round RESET + SEED commands; every
field of every battle after every step (tests/parity.py holds the comparison and the bar: integers bit-exact, fp32 with ==).

    python tools/fuzz_kernel_vs_oracle.py --backend gpu  [minutes]   # the sm_100a kernels through the C ABI (on a B200)
    python tools/fuzz_kernel_vs_oracle.py --backend host [minutes]   # csrc/frame_logic.cuh compiled for the host
                                                                     # (tests/host_emulation; GPU-less authoring container)
    python tools/fuzz_kernel_vs_oracle.py --backend gpu --engine ref # judged by oracle/_ref (the transliterated C#) instead

TEST INFRASTRUCTURE: the oracle is only ever the checker here."""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import parity_cases as pc                                   # noqa: E402
from parity import compare_state_and_outputs, compare_states, compare_stats, _eq  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("minutes", nargs="?", type=float, default=3.0)
ap.add_argument("--backend", choices=["gpu", "host"], default="gpu")
ap.add_argument("--engine", choices=["oracle", "ref"], default="oracle")
ap.add_argument("--master-seed", type=int, default=20261019)
args = ap.parse_args()

if args.engine == "ref":
    import ref_binding as eng
    Batch = eng.RefBatch
else:
    import oracle_binding as eng
    Batch = eng.OracleBatch

if args.backend == "gpu":
    from footsies_gym_b200 import FootsiesEnv
    if not torch.cuda.is_available():
        raise SystemExit("--backend gpu needs a CUDA device (this library has no CPU path)")

    def make_env(**kw):
        return FootsiesEnv(device="cuda:0", **kw)
else:
    from kernel_host import HostKernelEnv

    def make_env(**kw):
        return HostKernelEnv(**kw)

master = np.random.default_rng(args.master_seed)
t_end = time.time() + 60 * args.minutes
total_frames = total_episodes = rounds = 0
print(f"# backend={args.backend} engine={args.engine} master_seed={args.master_seed}", flush=True)
while time.time() < t_end:
    seed = int(master.integers(0, 2 ** 31 - 1))
    rng = np.random.default_rng(seed)
    p1_bot, p2_bot = bool(rng.random() < 0.2), bool(rng.random() < 0.6)
    cfg = dict(p1_bot=p1_bot, p2_bot=p2_bot, dense=bool(rng.random() < 0.7), autoreset=bool(rng.random() < 0.85),
               stale=bool(rng.random() < 0.9), first_env_index=int(rng.integers(0, 10 ** 6)),
               seed=int(rng.integers(-10 ** 6, 10 ** 6)), frame_skip=int(rng.choice([1, 1, 1, 2, 3, 4, 7])))
    n = int(rng.choice([33, 64, 255, 257, 300, 512, 777, 1024]))        # ragged tails and whole chunks of 256
    steps = int(rng.choice([300, 800, 1500]))
    # several chunks per pipeline group + a ragged tail (not with --engine ref: a transliterated game object graph is ~0.6 MB per battle)
    if rng.random() < 0.08 and args.engine != "ref":
        n, steps = int(rng.integers(120000, 260000)), 60
    resets = bool(rng.random() < 0.4)                                    # masked RESET + SEED commands in mid-round
    # FootsiesEnv frame_delay (footsies.py:129-131, 533-535: observation and info arrive `delay` steps late, reward and
    # termination do not): the delay ring kernel, CUDA path only (the host-compiled logic has no ring)
    # (only with frame_skip = 1, the reference's case: with K fused frames per step FootsiesEnv counts the delay in env steps,
    # while the oracle's repeat = K is K reference steps, each of which pushes to the queue)
    delay = int(rng.choice([0, 0, 0, 1, 3, 7])) if args.backend == "gpu" and n <= 1024 and cfg["frame_skip"] == 1 else 0
    maker = [pc.tape_uniform, pc.tape_sticky, pc.tape_profiles][int(rng.integers(0, 3))]
    t1, t2 = maker(rng, steps, n), maker(rng, steps, n)
    env = make_env(num_envs=n, by_example=p1_bot, opponent=None if p2_bot else "self_play", dense_reward=cfg["dense"],
                   frame_skip=cfg["frame_skip"], autoreset=cfg["autoreset"], seed=cfg["seed"],
                   first_env_index=cfg["first_env_index"], stale_intro_input=cfg["stale"], **({"frame_delay": delay} if delay else {}))
    orc = Batch(n, p1_bot=p1_bot, p2_bot=p2_bot, dense_reward=cfg["dense"], autoreset=cfg["autoreset"], frame_delay=delay,
                stale_intro_input=cfg["stale"], first_env_index=cfg["first_env_index"], seed=cfg["seed"], threads=8)
    cfg["frame_delay"] = delay

    def check(ret, where, with_reward=True):
        """delay = 0: every field of state and outputs; delay > 0: the battle state as it is now, the DELAYED observation and
        info the call returned, the undelayed reward and termination."""
        if not delay:
            compare_state_and_outputs(env, orc.trace, where=where)
            return
        compare_states(env.get_state(), orc.trace, where)
        obs, info = (ret[0], ret[-1])
        _eq(where, "delayed obs", torch.cat([obs[k].float() for k in ("guard", "move", "move_frame", "position")], 1).cpu().numpy(), orc.trace["obs"])
        _eq(where, "delayed info frame", info["frame"].cpu().numpy(), orc.trace["info_frame"])
        if with_reward:
            _eq(where, "reward", ret[1].cpu().numpy(), orc.trace["reward"])
            _eq(where, "terminated", ret[2].cpu().numpy().astype(np.int32), orc.trace["terminated"])

    ret = env.reset()
    orc.reset()
    where = f"seed {seed} {cfg} n={n}"
    check(ret, where + " reset", with_reward=False)
    for t in range(steps):
        if resets and rng.random() < 0.01:
            mask = rng.random(n) < 0.3
            new_seed = int(rng.integers(-10 ** 6, 10 ** 6)) if rng.random() < 0.5 else None
            ret = env.reset(seed=new_seed, options={"mask": torch.from_numpy(mask) if args.backend == "gpu" else mask})
            if new_seed is not None:
                orc.seed(new_seed, mask)
            orc.reset(mask)
            check(ret, where + f" masked reset before step {t}", with_reward=False)
        a1 = None if p1_bot else t1[t]
        a2 = None if p2_bot else t2[t]
        ret = env.step(None if a1 is None else torch.from_numpy(a1), None if a2 is None else torch.from_numpy(a2))
        orc.step(a1 if a1 is not None else np.zeros(n, np.uint8), a2, repeat=cfg["frame_skip"])
        check(ret, where + f" step {t}")
    if not resets:
        compare_stats(env, orc, where=where + " end")
    st = env.episode_stats()
    env.close()
    del orc
    rounds += 1
    total_frames += st["env_frames"]
    total_episodes += st["episodes"]
    print(f"round {rounds}: seed {seed} n={n} steps={steps} resets={resets} {cfg} -> {st['env_frames']} frames, "
          f"{st['episodes']} episodes, {st['guard_breaks']} guard breaks, {st['double_ko']} double KOs: identical", flush=True)
print(f"TOTAL: {rounds} random configurations, {total_frames} frames, {total_episodes} episodes: every field of every battle "
      f"identical between the {'CUDA path (C ABI)' if args.backend == 'gpu' else 'host-compiled kernel logic'} and "
      f"{'oracle/_ref (transliterated reference)' if args.engine == 'ref' else 'oracle/'} after every step")
