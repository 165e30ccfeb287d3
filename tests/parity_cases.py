"""Parity cases shared by the GPU tests (CUDA path through the C ABI) and the host-emulation tests (the same device
logic compiled for the host): the implementation under test against the CPU oracle on the same seeded action tapes
-- every env, every frame, every field (see tests/parity.py for the bar).  `scale` shrinks env counts for the CPU run."""
import numpy as np
import torch


def tape_uniform(rng, steps, n):
    """iid uniform over the 8 input bitmasks (the benchmark's synthetic actions)."""
    return rng.integers(0, 8, size=(steps, n), dtype=np.uint8)


def tape_sticky(rng, steps, n, p_change=0.15, weights=None):
    """Inputs held for geometric durations: reaches charged specials, dashes, long blocks, guard breaks."""
    out = np.zeros((steps, n), dtype=np.uint8)
    cur = rng.integers(0, 8, size=n, dtype=np.uint8)
    w = None if weights is None else np.asarray(weights, dtype=np.float64) / np.sum(weights)
    for t in range(steps):
        change = rng.random(n) < p_change
        new = rng.choice(8, size=n, p=w).astype(np.uint8)
        cur = np.where(change, new, cur)
        out[t] = cur
    return out


def run_case(make_env, ob, n, steps, *, p1_bot=False, p2_bot=True, dense=True, frame_skip=1, autoreset=True, stale=True,
             seed=0, tape1=None, tape2=None, first_env_index=0, check_every=1):
    from parity import compare_state_and_outputs, compare_stats
    env = make_env(num_envs=n, by_example=p1_bot, opponent=None if p2_bot else "self_play",
                   dense_reward=dense, frame_skip=frame_skip, autoreset=autoreset, seed=seed,
                   first_env_index=first_env_index, stale_intro_input=stale)
    orc = ob.OracleBatch(n, p1_bot=p1_bot, p2_bot=p2_bot, dense_reward=dense, autoreset=autoreset,
                         stale_intro_input=stale, first_env_index=first_env_index, seed=seed, threads=8)
    env.reset()
    orc.reset()
    compare_state_and_outputs(env, orc.trace, where="reset")
    for t in range(steps):
        a1 = None if p1_bot else tape1[t]
        a2 = None if p2_bot else tape2[t]
        env.step(None if a1 is None else torch.from_numpy(a1), None if a2 is None else torch.from_numpy(a2))
        orc.step(a1 if a1 is not None else np.zeros(n, np.uint8), a2, repeat=frame_skip)
        if t % check_every == 0 or t == steps - 1:
            compare_state_and_outputs(env, orc.trace, where=f"step {t}")
    compare_stats(env, orc, where="end")
    st = env.episode_stats()
    env.close()
    return st


def case_config_b_4096_envs_random_vs_bot_every_frame(make_env, oracle, scale=1.0):
    """BASELINE.json configs[1]: 4096 envs, random P1 vs BattleAI, frame-skip 1, 2048 frames, all compared."""
    rng = np.random.default_rng(1234)
    n, steps = max(64, int(4096 * scale)), 2048
    st = run_case(make_env, oracle, n, steps, tape1=tape_uniform(rng, steps, n))
    assert st["episodes"] > 1000 * scale and st["hits"] > 0 and st["blocks"] > 0


def case_self_play_sticky_inputs(make_env, oracle, scale=1.0):
    rng = np.random.default_rng(7)
    n, steps = max(64, int(2048 * scale)), 1500
    st = run_case(make_env, oracle, n, steps, p2_bot=False,
                  tape1=tape_sticky(rng, steps, n), tape2=tape_sticky(rng, steps, n, p_change=0.1))
    assert st["episodes"] > 100 * scale and st["guard_breaks"] > 0 and st["p1_specials"] > 0 and st["double_ko"] >= 0


def case_self_play_blockers_reach_guard_break_and_proximity(make_env, oracle, scale=1.0):
    rng = np.random.default_rng(11)
    n, steps = max(64, int(1024 * scale)), 1500
    # P2 mostly holds back (Right = 2) -> blocks, proximity guard, guard breaks
    st = run_case(make_env, oracle, n, steps, p2_bot=False,
                  tape1=tape_sticky(rng, steps, n, weights=[1, 0.2, 3, 0.2, 2, 0.2, 3, 0.2]),
                  tape2=tape_sticky(rng, steps, n, weights=[1, 0.5, 6, 0.2, 1, 0.2, 1, 0.1]))
    assert st["guard_breaks"] > 10 * scale and st["blocks"] > 100 * scale


def case_both_bots_by_example(make_env, oracle, scale=1.0):
    st = run_case(make_env, oracle, max(64, int(1024 * scale)), 1500, p1_bot=True, p2_bot=True, seed=99)
    assert st["episodes"] > 50 * scale


def case_p1_bot_vs_remote_p2(make_env, oracle, scale=1.0):
    rng = np.random.default_rng(5)
    n, steps = max(64, int(512 * scale)), 1000
    run_case(make_env, oracle, n, steps, p1_bot=True, p2_bot=False, tape2=tape_sticky(rng, steps, n), seed=3)


def case_sparse_reward_and_global_index_offset(make_env, oracle, scale=1.0):
    rng = np.random.default_rng(2)
    n, steps = max(64, int(1024 * scale)), 1000
    run_case(make_env, oracle, n, steps, dense=False, tape1=tape_sticky(rng, steps, n), first_env_index=123456, seed=-5)


def case_autoreset_disabled_freezes_done_envs(make_env, oracle, scale=1.0):
    rng = np.random.default_rng(3)
    n, steps = max(64, int(512 * scale)), 1200
    st = run_case(make_env, oracle, n, steps, autoreset=False, tape1=tape_uniform(rng, steps, n))
    assert st["episodes"] <= n and st["episodes"] > n // 4


def case_stale_intro_input_off(make_env, oracle, scale=1.0):
    rng = np.random.default_rng(4)
    n, steps = max(64, int(512 * scale)), 1000
    run_case(make_env, oracle, n, steps, stale=False, p2_bot=False, tape1=tape_sticky(rng, steps, n),
             tape2=tape_sticky(rng, steps, n))


def case_fused_frame_skip(make_env, oracle, k, p2_bot, scale=1.0):
    """configs[2]: self-play, K = 4 fused per launch (plus bot / odd K variants).  K = 2 runs the input personalities,
    which reach charged specials, dashes and guard breaks under frame skip (the fused kernels decide requests through
    the request lookup table, the single-frame kernels through the select chain: both are covered)."""
    rng = np.random.default_rng(100 + k)
    n, steps = max(64, int(4096 * scale)), 400 if k != 2 else 1200
    tape = (lambda: tape_profiles(rng, steps, n)) if k == 2 else (lambda: tape_sticky(rng, steps, n, p_change=0.3))
    st = run_case(make_env, oracle, n, steps, p2_bot=p2_bot, frame_skip=k, tape1=tape(), tape2=None if p2_bot else tape())
    assert st["episodes"] > 100 * scale
    if k == 2:
        assert st["p1_specials_neutral"] > 0 and st["guard_breaks"] > 0


def tape_profiles(rng, steps, n):
    """Every env gets its own input personality: how long it holds an input and which of the 8 combinations it
    prefers (turtles, dash spammers, button holders, ...).  Reaches corners of the state space that uniform or
    uniformly sticky tapes visit rarely: long charges released at odd moments, double KOs, guard breaks in the corner."""
    p_change = rng.choice([0.02, 0.05, 0.15, 0.4, 0.9], size=n)
    prefs = rng.dirichlet(np.full(8, 0.35), size=n)                      # sparse preferences per env
    cdf = np.cumsum(prefs, axis=1)
    cur = (rng.random(n)[:, None] > cdf).sum(axis=1).astype(np.uint8)
    out = np.zeros((steps, n), dtype=np.uint8)
    for t in range(steps):
        change = rng.random(n) < p_change
        new = np.minimum((rng.random(n)[:, None] > cdf).sum(axis=1), 7).astype(np.uint8)
        cur = np.where(change, new, cur).astype(np.uint8)
        out[t] = cur
    return out


def case_input_personalities_self_play(make_env, oracle, scale=1.0):
    rng = np.random.default_rng(2718)
    n, steps = max(64, int(4096 * scale)), 2500
    st = run_case(make_env, oracle, n, steps, p2_bot=False, tape1=tape_profiles(rng, steps, n),
                  tape2=tape_profiles(rng, steps, n))
    assert st["episodes"] > 200 * scale and st["guard_breaks"] > 0 and st["p1_specials_neutral"] > 0


def case_input_personalities_vs_bot_sparse(make_env, oracle, scale=1.0):
    rng = np.random.default_rng(3141)
    n, steps = max(64, int(4096 * scale)), 2500
    st = run_case(make_env, oracle, n, steps, dense=False, tape1=tape_profiles(rng, steps, n), seed=77)
    assert st["episodes"] > 200 * scale and st["p1_specials_neutral"] > 0


FUSED_PARAMS = [(4, False), (4, True), (3, True), (16, False), (2, False), (2, True)]


def _obs_is_skippable(obs):
    """FootsiesFrameSkipped._is_obs_skippable (wrappers/frame_skip.py:56-66) on the obs tensor [N, 8]."""
    o = np.asarray(obs.cpu() if hasattr(obs, "cpu") else obs)
    m1, m2, mf1 = o[:, 2].astype(np.int64), o[:, 3].astype(np.int64), o[:, 4]
    hit_guard = np.zeros(17, dtype=bool)
    hit_guard[[9, 10, 11, 12, 13]] = True          # DAMAGE, GUARD_M, GUARD_STAND, GUARD_CROUCH, GUARD_BREAK (moves.py order)
    return ((mf1 != 0.0) & ~hit_guard[m2]) | (m1 == 9)


def frame_skipped_fused_vs_masked_loop(make_env, frame_skip=1, p2_bot=True, n=400, steps=300, seed=3, dense=True,
                                       autoreset=True):
    """fg_config.skip_unactionable (FootsiesFrameSkipped fused into the step) against the wrapper's own loop -- one env step,
    then masked no-op steps for the envs whose observation P1 cannot act on (wrappers/frame_skip.py:68-80): identical
    state, observation, termination and statistics after every wrapper step; summed reward to 1e-6 (the loop adds float32
    step rewards, the kernel sums in float64 like the reference's Python floats)."""
    rng = np.random.default_rng(seed)
    kw = dict(num_envs=n, opponent=None if p2_bot else "self_play", frame_skip=frame_skip, seed=seed, dense_reward=dense,
              autoreset=autoreset)
    fused, plain = make_env(**kw), make_env(**kw)
    fused.set_skip_unactionable(True)
    fused.reset()
    plain.reset()
    tape1 = tape_sticky(rng, steps, n, p_change=0.3)
    tape2 = tape_sticky(rng, steps, n, p_change=0.3)
    zero = np.zeros(n, dtype=np.uint8)
    skipped_steps = 0
    for t in range(steps):
        a2 = None if p2_bot else tape2[t]
        fused.step(tape1[t], a2)
        plain.step(tape1[t], a2)
        total = np.asarray(plain.reward.cpu(), dtype=np.float64).copy()
        done = np.asarray(plain.terminated.cpu()).astype(bool)
        skip = _obs_is_skippable(plain.obs) & ~done
        while skip.any():
            plain.set_step_mask(torch.from_numpy(skip))
            plain.step(zero, a2)
            plain.set_step_mask(None)
            total += np.where(skip, np.asarray(plain.reward.cpu(), dtype=np.float64), 0.0)
            done = np.asarray(plain.terminated.cpu()).astype(bool)
            skip = skip & _obs_is_skippable(plain.obs) & ~done
            skipped_steps += 1
        where = f"frame-skipped step {t}"
        assert fused.get_state().tobytes() == plain.get_state().tobytes(), where
        assert np.array_equal(np.asarray(fused.obs.cpu()), np.asarray(plain.obs.cpu())), where
        assert np.array_equal(np.asarray(fused.terminated.cpu()).astype(bool), done), where
        assert np.array_equal(np.asarray(fused.info_frame.cpu()), np.asarray(plain.info_frame.cpu())), where
        assert np.abs(np.asarray(fused.reward.cpu(), dtype=np.float64) - total).max() <= 1e-6, where
        assert not (_obs_is_skippable(fused.obs) & ~done).any(), where
    assert skipped_steps > (steps if autoreset else 10)     # the loop really had work to do
    assert fused.episode_stats() == plain.episode_stats()
    fused.close()
    plain.close()
