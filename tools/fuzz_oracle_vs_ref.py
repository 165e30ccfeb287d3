#!/usr/bin/env python3
"""Differential fuzzing of the two CPU engines beyond the fixed tapes of the test-suite: the hand-written oracle
(oracle/footsies_oracle.c) against the transliterated reference (oracle/_ref, tools/cs2cpp.py) on randomly drawn
configurations, seeds and input personalities, every field of every trace after every step (tests/test_oracle_vs_ref.py
holds the comparison).  Usage: python tools/fuzz_oracle_vs_ref.py [minutes] [master seed] > profiles/rNN_oracle_vs_ref_fuzz.log"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import oracle_binding as ob   # noqa: E402
import parity_cases as pc     # noqa: E402
import ref_binding as rb      # noqa: E402
from test_oracle_vs_ref import assert_traces_equal   # noqa: E402

minutes = float(sys.argv[1]) if len(sys.argv) > 1 else 5.0
master_seed = int(sys.argv[2]) if len(sys.argv) > 2 else 20261018
print(f"# master_seed={master_seed}", flush=True)
master = np.random.default_rng(master_seed)
t_end = time.time() + 60 * minutes
total_frames = total_episodes = rounds = 0
while time.time() < t_end:
    seed = int(master.integers(0, 2 ** 31 - 1))
    rng = np.random.default_rng(seed)
    cfg = dict(p1_bot=bool(rng.random() < 0.25), p2_bot=bool(rng.random() < 0.6), dense_reward=bool(rng.random() < 0.7),
               frame_delay=int(rng.choice([0, 0, 1, 3, 7])), autoreset=bool(rng.random() < 0.85),
               stale_intro_input=bool(rng.random() < 0.9), first_env_index=int(rng.integers(0, 10 ** 6)),
               seed=int(rng.integers(-10 ** 6, 10 ** 6)))
    n, steps, repeat = int(rng.choice([32, 64, 128])), int(rng.choice([600, 1500, 3000])), int(rng.choice([1, 1, 1, 2, 4]))
    o, r = ob.OracleBatch(n, threads=8, **cfg), rb.RefBatch(n, threads=8, **cfg)
    o.reset()
    r.reset()
    assert_traces_equal(o.trace, r.trace, f"seed {seed} reset")
    maker = [pc.tape_uniform, pc.tape_sticky, pc.tape_profiles][int(rng.integers(0, 3))]
    t1, t2 = maker(rng, steps, n), maker(rng, steps, n)
    for t in range(steps):
        a2 = None if cfg["p2_bot"] else t2[t]
        o.step(t1[t], a2, repeat=repeat)
        r.step(t1[t], a2, repeat=repeat)
        assert_traces_equal(o.trace, r.trace, f"seed {seed} cfg {cfg} step {t}")
        if rng.random() < 0.002:                       # RESET + SEED in mid-round on a random subset
            mask = rng.random(n) < 0.3
            s = int(rng.integers(0, 10 ** 6))
            for b in (o, r):
                b.seed(s, mask)
                b.reset(mask)
            assert_traces_equal(o.trace, r.trace, f"seed {seed} masked reset at {t}")
    st = o.stats()
    assert st == r.stats()
    total_frames += o.frames_simulated()
    total_episodes += st["episodes"]
    rounds += 1
    print(f"round {rounds}: seed {seed} n={n} steps={steps} repeat={repeat} {cfg} -> {o.frames_simulated()} frames, "
          f"{st['episodes']} episodes, {st['guard_breaks']} guard breaks, {st['double_ko']} double KOs: identical", flush=True)
print(f"TOTAL: {rounds} random configurations, {total_frames} frames, {total_episodes} episodes: every field of every trace "
      f"byte-identical between oracle/ and oracle/_ref")
