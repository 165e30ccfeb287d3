#!/usr/bin/env python3
"""Small deterministic workload for compute-sanitizer (memcheck / racecheck / synccheck) runs of the step kernel:
ragged batch (tail chunk), bot + self-play + fused + masked variants, a few hundred steps, checked against the oracle.

    compute-sanitizer --tool racecheck python tools/sanitize_check.py
Build a single-shape library first to force the large CTA shapes at this small size, e.g.
    tools/probes/build_variant.sh san768 -DFG_THREADS=768 -DFG_GROUP=256 -DFG_STAGES=3 -DFG_BLOCKS_PER_SM=1
    FOOTSIES_B200_LIB=$PWD/tools/probes/lib_san768.so compute-sanitizer ... python tools/sanitize_check.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch

import oracle_binding as ob
from footsies_gym_b200 import FootsiesEnv
from parity import compare_state_and_outputs

steps = int(os.environ.get("SAN_STEPS", "60"))
rng = np.random.default_rng(0)
# The last two batches give every pipeline group SEVERAL chunks plus a ragged tail (230 000 = 898 x 256 + 112 battles over at most
# 444 / 592 groups): consecutive launches walk the chunks in alternating directions, and the tail chunk must stay the final
# iteration of its group in both (fewer steps: the oracle follows every battle).
BIG = int(os.environ.get("SAN_BIG", "230000"))
for (n, p2_bot, k) in ((2000, True, 1), (1300, False, 1), (1800, True, 3), (900, False, 4), (BIG, True, 1), (BIG + 77, False, 2)):
    env = FootsiesEnv(num_envs=n, device="cuda:0", opponent=None if p2_bot else "self_play", frame_skip=k, seed=1)
    orc = ob.OracleBatch(n, p2_bot=p2_bot, seed=1)
    env.reset()
    orc.reset()
    for t in range(steps if n < 100000 else min(steps, 40)):
        a1 = rng.integers(0, 8, size=n, dtype=np.uint8)
        a2 = rng.integers(0, 8, size=n, dtype=np.uint8)
        env.step(torch.from_numpy(a1), None if p2_bot else torch.from_numpy(a2))
        orc.step(a1, None if p2_bot else a2, repeat=k)
        if t % 20 == 19:
            compare_state_and_outputs(env, orc.trace, where=f"n={n} step {t}")
    # masked stepping
    mask = torch.from_numpy(rng.random(n) < 0.5)
    env.set_step_mask(mask)
    env.step(torch.from_numpy(a1), None if p2_bot else torch.from_numpy(a2))
    env.set_step_mask(None)
    torch.cuda.synchronize()
    env.close()
    print("ok", n, p2_bot, k, flush=True)
print("sanitize_check done")
