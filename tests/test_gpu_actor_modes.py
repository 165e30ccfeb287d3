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
