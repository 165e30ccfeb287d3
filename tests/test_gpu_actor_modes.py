"""SURVEY.md §8f row 4 and row a18 on the GPU: python-callable opponent, runtime set_opponent toggle (P2_BOT command,
footsies.py:458-480), and the frame_delay queue (footsies.py:129-131, 502-504, 533-535) for a whole batch with
automatic restarts."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _cuda():
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible")
    return torch.device("cuda:0")


def test_callable_opponent_is_queried_every_step_with_the_latest_observation(oracle):
    from footsies_gym_b200 import FootsiesEnv
    dev = _cuda()
    n, steps = 300, 400
    rng = np.random.default_rng(21)
    tape1 = rng.integers(0, 8, size=(steps, n), dtype=np.uint8)
    seen = []

    def opponent(obs, info):
        # like the reference's opponent callable (footsies.py:522-527): sees what the agent saw last
        seen.append((obs["position"].clone(), info["frame"].clone()))
        # a reactive policy: hold back (Right for P2) when P1 is close, else walk forward (Left); attack on even frames
        close = (obs["position"][:, 1] - obs["position"][:, 0]) < 1.8
        attack = (info["frame"] % 2 == 0).to(torch.uint8) * 4
        return torch.where(close, torch.tensor(2, dtype=torch.uint8, device=dev), torch.tensor(1, dtype=torch.uint8, device=dev)) | attack

    env = FootsiesEnv(num_envs=n, device=dev, opponent=opponent, seed=0)
    orc = oracle.OracleBatch(n, p2_bot=False, seed=0)
    env.reset()
    orc.reset()
    for t in range(steps):
        prev_pos = orc.trace["obs"][:, 6:8].copy()
        prev_frame = orc.trace["info_frame"].copy()
        a2 = np.where((prev_pos[:, 1] - prev_pos[:, 0]) < 1.8, 2, 1).astype(np.uint8) | ((prev_frame % 2 == 0).astype(np.uint8) * 4)
        env.step(torch.from_numpy(tape1[t]))
        orc.step(tape1[t], a2)
        assert np.array_equal(env.obs.cpu().numpy(), orc.trace["obs"]), t
        assert np.array_equal(env.reward.cpu().numpy(), orc.trace["reward"]), t
        assert np.array_equal(seen[-1][0].cpu().numpy(), prev_pos) and np.array_equal(seen[-1][1].cpu().numpy(), prev_frame)
    assert len(seen) == steps
    env.close()


def test_set_opponent_toggles_between_policy_and_in_game_bot(oracle):
    from footsies_gym_b200 import FootsiesEnv
    dev = _cuda()
    n = 256
    rng = np.random.default_rng(22)
    env = FootsiesEnv(num_envs=n, device=dev, opponent=lambda obs, info: torch.zeros(n, dtype=torch.uint8, device=dev), seed=5)
    env.reset()
    for t in range(50):
        # both idle: nothing of this phase (held inputs, hit stun, guard flags) leaks into the next round, so the
        # bot phase below can be compared with a fresh oracle
        env.step(torch.zeros(n, dtype=torch.uint8))
    assert int(env.info_frame.min()) == 49
    assert env.set_opponent(None) is True                      # back to the in-game bot; reset afterwards (footsies.py:466)
    env.reset(seed=5)
    orc = oracle.OracleBatch(n, p2_bot=True, seed=5)
    orc.reset()
    for t in range(300):
        a = rng.integers(0, 8, size=n, dtype=np.uint8)
        env.step(torch.from_numpy(a))
        orc.step(a)
        assert np.array_equal(env.obs.cpu().numpy(), orc.trace["obs"]), t
        assert np.array_equal(env.terminated.cpu().numpy().astype(np.int32), orc.trace["terminated"]), t
    # and to a policy again: P2 idles, so it never attacks
    env.set_opponent(lambda obs, info: torch.zeros(n, dtype=torch.uint8, device=dev))
    env.reset()
    for t in range(100):
        env.step(torch.from_numpy(rng.integers(0, 8, size=n, dtype=np.uint8)))
        assert int(env.info_misc[:, 1].max()) == 0             # p2_action mask stays empty
    env.close()


@pytest.mark.parametrize("delay", [1, 4])
def test_frame_delay_queue_for_a_batch_with_automatic_restarts(oracle, delay):
    from footsies_gym_b200 import FootsiesEnv
    dev = _cuda()
    n, steps = 512, 900
    rng = np.random.default_rng(30 + delay)
    env = FootsiesEnv(num_envs=n, device=dev, frame_delay=delay, seed=2)
    orc = oracle.OracleBatch(n, p2_bot=True, frame_delay=delay, seed=2)
    obs, info = env.reset()
    orc.reset()
    assert np.array_equal(torch.cat([obs[k] for k in ("guard", "move", "move_frame", "position")], 1).cpu().numpy(), orc.trace["obs"])
    episodes = 0
    for t in range(steps):
        a = rng.integers(0, 8, size=n, dtype=np.uint8)
        obs, reward, term, trunc, info = env.step(torch.from_numpy(a))
        orc.step(a)
        got = torch.cat([obs[k] for k in ("guard", "move", "move_frame", "position")], 1).cpu().numpy()
        assert np.array_equal(got, orc.trace["obs"]), (t, np.argwhere(got != orc.trace["obs"])[:3])
        assert np.array_equal(info["frame"].cpu().numpy(), orc.trace["info_frame"]), t
        assert np.array_equal(reward.cpu().numpy(), orc.trace["reward"]), t          # reward / termination are not delayed
        assert np.array_equal(term.cpu().numpy().astype(np.int32), orc.trace["terminated"]), t
        episodes += int(term.sum())
    assert episodes > n
    env.close()


def test_frame_delay_on_the_host_buffer_path(oracle):
    """step_host / reset_host with frame_delay: the delayed observation and info reach the pinned host tensors (compact
    layout), reward and termination undelayed."""
    from footsies_gym_b200 import FootsiesEnv
    dev = _cuda()
    n, steps, delay = 300, 400, 3
    rng = np.random.default_rng(77)
    env = FootsiesEnv(num_envs=n, device=dev, frame_delay=delay, seed=4)
    orc = oracle.OracleBatch(n, p2_bot=True, frame_delay=delay, seed=4)
    keys = ("guard", "move", "move_frame", "position")
    obs, info = env.reset_host()
    orc.reset()
    assert not obs["position"].is_cuda and obs["guard"].dtype == torch.uint8
    assert np.array_equal(np.concatenate([obs[k].numpy() for k in keys], 1), orc.trace["obs"])
    for t in range(steps):
        a = rng.integers(0, 8, size=n, dtype=np.uint8)
        obs, reward, term, trunc, info = env.step_host(a)
        orc.step(a)
        assert np.array_equal(np.concatenate([obs[k].numpy() for k in keys], 1), orc.trace["obs"]), t
        assert np.array_equal(info["frame"].numpy(), orc.trace["info_frame"]), t
        assert np.array_equal(np.stack([info[k].numpy() for k in ("p1_action", "p2_action")], 1), orc.trace["info_action"]), t
        assert np.array_equal(reward.numpy(), orc.trace["reward"]), t
        assert np.array_equal(term.numpy().astype(np.int32), orc.trace["terminated"]), t
    env.close()


@pytest.mark.parametrize("autoreset", [False, True])
def test_frame_skipped_wrapper_over_a_delayed_batch(oracle, autoreset):
    """FootsiesFrameSkipped(FootsiesEnv(num_envs > 1, frame_delay = 3)): the wrapper's loop of masked steps must leave the
    frame_delay queue of the battles it holds back untouched (the reference keeps one deque per env, footsies.py:129-131,
    533-535).  Checked against one CPU oracle per battle driven by the reference wrapper's own loop
    (wrappers/frame_skip.py:68-80) on the DELAYED observation."""
    from footsies_gym_b200 import FootsiesEnv
    from footsies_gym_b200.wrappers import FootsiesFrameSkipped
    import parity_cases as pc
    dev = _cuda()
    n, steps, delay = 96, 260, 3
    rng = np.random.default_rng(41)
    env = FootsiesFrameSkipped(FootsiesEnv(num_envs=n, device=dev, frame_delay=delay, seed=6, autoreset=autoreset))
    assert not env.fused
    orcs = [oracle.OracleBatch(1, p2_bot=True, frame_delay=delay, seed=6, first_env_index=i, autoreset=autoreset) for i in range(n)]
    obs, info = env.reset()
    for o in orcs:
        o.reset()
    tape = pc.tape_sticky(rng, steps, n, p_change=0.3)
    zero = np.zeros(1, np.uint8)
    held_back = 0
    for t in range(steps):
        obs, reward, term, trunc, info = env.step(torch.from_numpy(tape[t]))
        exp_obs, exp_rew, exp_term, exp_frame = [], [], [], []
        for i, o in enumerate(orcs):
            if o.trace["terminated"][0] and not autoreset:     # autoreset off: a finished battle stays as it is
                tr, total = o.trace, 0.0
            else:
                tr = o.step(tape[t, i:i + 1])
                total = float(tr["reward_f64"][0])
                while pc._obs_is_skippable(tr["obs"])[0] and not tr["terminated"][0]:
                    tr = o.step(zero)
                    total += float(tr["reward_f64"][0])
                    held_back += 1
            exp_obs.append(tr["obs"][0].copy()); exp_rew.append(total)
            exp_term.append(int(tr["terminated"][0])); exp_frame.append(int(tr["info_frame"][0]))
        exp_obs = np.stack(exp_obs)
        got = torch.cat([obs["guard"], obs["move"], obs["move_frame"], obs["position"]], 1).cpu().numpy()
        assert np.array_equal(got, np.delete(exp_obs, 4, axis=1)), (t, np.argwhere(got != np.delete(exp_obs, 4, axis=1))[:3])
        assert np.array_equal(info["frame"].cpu().numpy(), np.asarray(exp_frame)), t
        assert np.array_equal(term.cpu().numpy().astype(np.int32), np.asarray(exp_term)), t
        # the wrapper adds float32 step rewards, the reference Python floats: 1e-6
        assert np.abs(reward.cpu().numpy().astype(np.float64) - np.asarray(exp_rew)).max() <= 1e-6, t
    assert held_back > steps
    env.close()


@pytest.mark.parametrize("dense,p2_bot", [(True, True), (False, True), (True, False)])
def test_packed_host_layout_is_lossless(oracle, monkeypatch, dense, p2_bot):
    """fg_step_host_packed / fg_reset_host_packed: one 16-byte record per battle (include/footsies_b200.h).  Decoding the
    records gives exactly what the CPU oracle says FootsiesEnv.step returns -- observation, reward (through the reward
    table), termination, info -- for a batch large enough to be cut into pipelined slices, masked resets included."""
    from footsies_gym_b200 import FootsiesEnv
    import parity_cases as pc
    dev = _cuda()
    n, steps = 3 * 1024 * 1024 // 8 + 77, 260
    monkeypatch.setenv("FOOTSIES_B200_HOST_CHUNK_ENVS", "131072")      # four pipelined slices, the last one ragged
    rng = np.random.default_rng(12)
    env = FootsiesEnv(num_envs=n, device=dev, dense_reward=dense, opponent=None if p2_bot else "self_play", seed=9)
    orc = oracle.OracleBatch(n, p2_bot=p2_bot, dense_reward=dense, seed=9, threads=8)
    keys = ("guard", "move", "move_frame", "position")

    def check(packed, where):
        assert packed.dtype == torch.int32 and tuple(packed.shape) == (n, 4) and not packed.is_cuda and packed.is_pinned()
        obs, reward, term, trunc, info = env.decode_packed(packed)
        got = np.concatenate([obs[k].numpy().astype(np.float32) for k in keys], 1)
        assert np.array_equal(got, orc.trace["obs"]), where
        assert np.array_equal(reward.numpy(), orc.trace["reward"]), where
        assert np.array_equal(term.numpy().astype(np.int32), orc.trace["terminated"]), where
        assert np.array_equal(info["frame"].numpy(), orc.trace["info_frame"]), where
        assert np.array_equal(np.stack([info[k].numpy() for k in ("p1_action", "p2_action")], 1), orc.trace["info_action"]), where
        assert np.array_equal(np.stack([info[k].numpy() for k in ("p1_hitstun", "p2_hitstun")], 1), orc.trace["info_hitstun"]), where
        assert not bool(trunc.any())
    orc.reset()
    check(env.reset_host_packed(), "reset")
    t1 = pc.tape_sticky(rng, steps, n, p_change=0.25)
    t2 = pc.tape_sticky(rng, steps, n, p_change=0.25)
    for t in range(steps):
        a2 = None if p2_bot else t2[t]
        orc.step(t1[t], a2)
        check(env.step_host_packed(t1[t], a2), f"step {t}")
        if t == 120:
            mask = rng.random(n) < 0.2
            orc.seed(33, mask)
            orc.reset(mask)
            check(env.reset_host_packed(seed=33, mask=mask), "masked reset")
    tab = env.packed_reward_table()
    assert tab.numel() == 128 and bool((tab[:10].diff() > 0).all())
    assert env.host_io_bytes_per_step(packed=True) == (n * (1 if p2_bot else 2), n * 16)
    env.close()


def test_packed_host_layout_refuses_what_it_cannot_carry():
    from footsies_gym_b200 import FootsiesEnv, _capi
    dev = _cuda()
    env = FootsiesEnv(num_envs=64, device=dev, frame_skip=4, seed=0)
    env.reset()
    with pytest.raises(_capi.FootsiesLibraryError):
        env.step_host_packed(np.zeros(64, np.uint8))
    env.close()
    env = FootsiesEnv(num_envs=64, device=dev, frame_delay=2, seed=0)
    env.reset()
    with pytest.raises(ValueError):
        env.step_host_packed(np.zeros(64, np.uint8))
    env.close()
