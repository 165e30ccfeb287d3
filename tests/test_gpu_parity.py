"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on the same seeded
action tapes -- every env, every frame, every field (see tests/parity.py for the bar)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

pytestmark = pytest.mark.gpu


def _require_cuda():
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible")


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


def run_case(ob, n, steps, *, p1_bot=False, p2_bot=True, dense=True, frame_skip=1, autoreset=True, stale=True,
             seed=0, tape1=None, tape2=None, first_env_index=0, check_every=1):
    from footsies_gym_b200 import FootsiesEnv
    from parity import compare_state_and_outputs, compare_stats
    _require_cuda()
    env = FootsiesEnv(num_envs=n, device="cuda:0", by_example=p1_bot, opponent=None if p2_bot else "self_play",
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


def test_config_b_4096_envs_random_vs_bot_every_frame(oracle):
    """BASELINE.json configs[1]: 4096 envs, random P1 vs BattleAI, frame-skip 1, 2048 frames, all compared."""
    rng = np.random.default_rng(1234)
    n, steps = 4096, 2048
    st = run_case(oracle, n, steps, tape1=tape_uniform(rng, steps, n))
    assert st["episodes"] > 1000 and st["hits"] > 0 and st["blocks"] > 0


def test_self_play_sticky_inputs(oracle):
    rng = np.random.default_rng(7)
    n, steps = 2048, 1500
    st = run_case(oracle, n, steps, p2_bot=False,
                  tape1=tape_sticky(rng, steps, n), tape2=tape_sticky(rng, steps, n, p_change=0.1))
    assert st["episodes"] > 100 and st["guard_breaks"] > 0 and st["p1_specials"] > 0 and st["double_ko"] >= 0


def test_self_play_blockers_reach_guard_break_and_proximity(oracle):
    rng = np.random.default_rng(11)
    n, steps = 1024, 1500
    # P2 mostly holds back (Right = 2) -> blocks, proximity guard, guard breaks
    st = run_case(oracle, n, steps, p2_bot=False,
                  tape1=tape_sticky(rng, steps, n, weights=[1, 0.2, 3, 0.2, 2, 0.2, 3, 0.2]),
                  tape2=tape_sticky(rng, steps, n, weights=[1, 0.5, 6, 0.2, 1, 0.2, 1, 0.1]))
    assert st["guard_breaks"] > 10 and st["blocks"] > 100


def test_both_bots_by_example(oracle):
    st = run_case(oracle, 1024, 1500, p1_bot=True, p2_bot=True, seed=99)
    assert st["episodes"] > 50


def test_p1_bot_vs_remote_p2(oracle):
    rng = np.random.default_rng(5)
    n, steps = 512, 1000
    run_case(oracle, n, steps, p1_bot=True, p2_bot=False, tape2=tape_sticky(rng, steps, n), seed=3)


def test_sparse_reward_and_global_index_offset(oracle):
    rng = np.random.default_rng(2)
    n, steps = 1024, 1000
    run_case(oracle, n, steps, dense=False, tape1=tape_sticky(rng, steps, n), first_env_index=123456, seed=-5)


def test_autoreset_disabled_freezes_done_envs(oracle):
    rng = np.random.default_rng(3)
    n, steps = 512, 1200
    st = run_case(oracle, n, steps, autoreset=False, tape1=tape_uniform(rng, steps, n))
    assert st["episodes"] <= n and st["episodes"] > n // 4


def test_stale_intro_input_off(oracle):
    rng = np.random.default_rng(4)
    n, steps = 512, 1000
    run_case(oracle, n, steps, stale=False, p2_bot=False, tape1=tape_sticky(rng, steps, n),
             tape2=tape_sticky(rng, steps, n))


@pytest.mark.parametrize("k,p2_bot", [(4, False), (4, True), (3, True), (16, False)])
def test_fused_frame_skip(oracle, k, p2_bot):
    """configs[2]: self-play, K = 4 fused per launch (plus bot / odd K variants)."""
    rng = np.random.default_rng(100 + k)
    n, steps = 4096, 400
    st = run_case(oracle, n, steps, p2_bot=p2_bot, frame_skip=k, tape1=tape_sticky(rng, steps, n, p_change=0.3),
                  tape2=None if p2_bot else tape_sticky(rng, steps, n, p_change=0.3))
    assert st["episodes"] > 100


def test_masked_hard_reset_and_reseed(oracle):
    from footsies_gym_b200 import FootsiesEnv
    from parity import compare_state_and_outputs
    _require_cuda()
    rng = np.random.default_rng(8)
    n = 777                                  # ragged: not a multiple of the warp or CTA size
    env = FootsiesEnv(num_envs=n, device="cuda:0", seed=1)
    orc = oracle.OracleBatch(n, p2_bot=True, seed=1)
    env.reset()
    orc.reset()
    for t in range(400):
        a = rng.integers(0, 8, size=n, dtype=np.uint8)
        env.step(torch.from_numpy(a))
        orc.step(a)
        if t % 50 == 25:
            mask = rng.random(n) < 0.3
            env.reset(seed=1000 + t, options={"mask": torch.from_numpy(mask)})
            orc.seed(1000 + t, mask)
            orc.reset(mask)
        compare_state_and_outputs(env, orc.trace, where=f"step {t}")
    env.close()


def test_single_env_reference_style_loop(oracle):
    """num_envs = 1 used exactly like the reference: tuple actions, reset() after termination."""
    from footsies_gym_b200 import FootsiesEnv
    _require_cuda()
    env = FootsiesEnv(autoreset=False, seed=0)
    orc = oracle.OracleBatch(1, p2_bot=True, autoreset=False, seed=0)
    rng = np.random.default_rng(0)
    episodes = 0
    for _ in range(3):
        obs, info = env.reset()
        tr = orc.reset()
        assert int(info["frame"][0]) == -1
        terminated = False
        while not terminated:
            a = tuple(bool(x) for x in rng.integers(0, 2, size=3))
            obs, reward, term, trunc, info = env.step(a)
            tr = orc.step([a[0] | a[1] << 1 | a[2] << 2])
            terminated = bool(term[0])
            assert terminated == bool(tr[0]["terminated"]) and not bool(trunc[0])
            assert float(reward[0]) == float(tr[0]["reward"])
            assert obs["position"][0].tolist() == tr[0]["obs"][6:8].tolist()
        episodes += 1
    assert episodes == 3
    env.close()


def test_host_buffer_step_matches_device_step(oracle):
    from footsies_gym_b200 import FootsiesEnv
    _require_cuda()
    n = 1000
    rng = np.random.default_rng(6)
    env = FootsiesEnv(num_envs=n, seed=5)
    orc = oracle.OracleBatch(n, p2_bot=True, seed=5)
    env.reset()
    orc.reset()
    for t in range(200):
        a = rng.integers(0, 8, size=n, dtype=np.uint8)
        obs, reward, term, trunc, info = env.step_host(a)
        tr = orc.step(a)
        assert np.array_equal(np.concatenate([obs[k].numpy() for k in ("guard", "move", "move_frame", "position")], 1),
                              tr["obs"])
        assert np.array_equal(reward.numpy(), tr["reward"])
        assert np.array_equal(term.numpy().astype(np.int32), tr["terminated"])
        assert np.array_equal(info["frame"].numpy(), tr["info_frame"])
    env.close()


def test_library_fails_loudly_on_bad_arguments():
    from footsies_gym_b200 import FootsiesEnv, _capi
    _require_cuda()
    env = FootsiesEnv(num_envs=4)
    with pytest.raises(RuntimeError):
        env.step(torch.zeros(4, dtype=torch.uint8))       # step before reset
    env.reset()
    with pytest.raises(ValueError):
        env.step(torch.zeros(5, dtype=torch.uint8))
    st = env.get_state()
    st["f"]["action_id"][0, 0] = 510                      # WIN is not representable
    with pytest.raises(_capi.FootsiesLibraryError):
        env.set_state(st)
    with pytest.raises(ValueError):
        FootsiesEnv(num_envs=4, render_mode="human")
    env.close()
