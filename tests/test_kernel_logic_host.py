"""The device frame logic (csrc/frame_logic.cuh -- the functions the sm_100a step kernel inlines) compiled for the
host by tests/host_emulation and run through the same parity cases as the GPU tests, at reduced env counts.

This is how kernel-logic changes are checked in the GPU-less authoring container before they go to a B200; it is test
infrastructure, not a CPU path of the product (which has none)."""
import numpy as np
import pytest

import parity_cases as pc


def make_env(**kw):
    from kernel_host import HostKernelEnv
    return HostKernelEnv(**kw)


CASES = [name for name in dir(pc) if name.startswith("case_") and name != "case_fused_frame_skip"]


@pytest.mark.parametrize("name", CASES, ids=[c[5:] for c in CASES])
def test_parity_case_host_logic(oracle, name):
    getattr(pc, name)(make_env, oracle, scale=0.125)


@pytest.mark.parametrize("k,p2_bot", pc.FUSED_PARAMS)
def test_fused_frame_skip_host_logic(oracle, k, p2_bot):
    pc.case_fused_frame_skip(make_env, oracle, k, p2_bot, scale=0.0625)


@pytest.mark.parametrize("frame_skip,p2_bot,dense,autoreset", [(1, True, True, True), (3, True, True, True),
                                                               (1, False, True, True), (2, True, False, True),
                                                               (1, True, True, False)])
def test_frame_skipped_fusion_host_logic(frame_skip, p2_bot, dense, autoreset):
    pc.frame_skipped_fused_vs_masked_loop(make_env, frame_skip, p2_bot, n=120, steps=250, dense=dense, autoreset=autoreset)


@pytest.mark.parametrize("by_example", [False, True])
def test_masked_reset_and_state_round_trip_host_logic(oracle, by_example):
    """by_example: P1's bot is never Reset() (spectator wrapper, BattleCore.cs:274) -- a RESET in mid-round makes it decide
    on the state it recorded last, a reset after a KO on the state before the terminal frame."""
    from kernel_host import HostKernelEnv
    from parity import compare_state_and_outputs
    rng = np.random.default_rng(8)
    n = 333
    env = HostKernelEnv(num_envs=n, seed=1, by_example=by_example)
    orc = oracle.OracleBatch(n, p1_bot=by_example, p2_bot=True, seed=1)
    env.reset()
    orc.reset()
    for t in range(400):
        a = rng.integers(0, 8, size=n, dtype=np.uint8)
        env.step(a)
        orc.step(a)
        if t % 50 == 25:
            mask = rng.random(n) < 0.3
            env.reset(seed=1000 + t, options={"mask": mask})
            orc.seed(1000 + t, mask)
            orc.reset(mask)
        if t % 40 == 7:
            env.set_state(env.get_state())           # decode -> encode must be the identity on reachable states
        compare_state_and_outputs(env, orc.trace, where=f"step {t}")
    env.close()
