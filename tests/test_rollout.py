"""BASELINE.json configs[4]: PPO-style rollout with a torch MLP policy reading the observation tensor in place.
The collected transitions must be exactly what a fresh env produces when the recorded actions are replayed into it
(and into the CPU oracle), both for the eager loop and for the CUDA-graph-captured horizon."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("use_graph,fused", [(False, False), (True, False), (False, "step"), (True, "step"),
                                             (False, "horizon")])
def test_rollout_is_replayable(oracle, use_graph, fused):
    from footsies_gym_b200 import FootsiesEnv
    from footsies_gym_b200.rollout import MLPPolicy, RolloutCollector
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible")
    n, horizon, rounds = 512, 64, 3
    torch.manual_seed(0)
    dev = torch.device("cuda:0")
    env = FootsiesEnv(num_envs=n, device=dev, seed=11)
    policy = MLPPolicy().to(dev)
    col = RolloutCollector(env, policy, horizon=horizon, use_cuda_graph=use_graph, fused=fused)
    # the graph path warms up and captures with real env steps, so the replay twin is driven from get_state snapshots
    ref = FootsiesEnv(num_envs=n, device=dev, seed=11)
    orc = oracle.OracleBatch(n, p2_bot=True, seed=11)
    col.collect()            # priming call: with use_cuda_graph it runs a warm-up horizon, captures, then replays
    for r in range(rounds):
        before = env.get_state()
        out = col.collect()
        torch.cuda.synchronize()
        ref.reset()
        ref.set_state(before)
        acts = out["actions"].cpu().numpy()
        obs = out["obs"].cpu().numpy()
        rew = out["rewards"].cpu().numpy()
        done = out["dones"].cpu().numpy()
        assert acts.max() <= 7 and len(np.unique(acts)) > 4                 # a stochastic policy over the 8 combinations
        for t in range(horizon):
            if t > 0:
                assert np.array_equal(obs[t], ref.obs.cpu().numpy()), (r, t)   # obs[t] is what the policy saw at step t
            ref.step(torch.from_numpy(acts[t]))
            assert np.array_equal(rew[t], ref.reward.cpu().numpy()), (r, t)
            assert np.array_equal(done[t], ref.terminated.cpu().numpy()), (r, t)
        assert np.array_equal(out["last_obs"].cpu().numpy(), ref.obs.cpu().numpy())
    # the oracle agrees with the replay twin on a full replay from reset of the first horizon's actions
    env2 = FootsiesEnv(num_envs=n, device=dev, seed=11)
    col2 = RolloutCollector(env2, policy, horizon=horizon, use_cuda_graph=False, fused=fused)
    out = col2.collect()
    orc.reset()
    for t in range(horizon):
        tr = orc.step(out["actions"][t].cpu().numpy())
        assert np.array_equal(out["rewards"][t].cpu().numpy(), tr["reward"])
        assert np.array_equal(out["dones"][t].cpu().numpy().astype(np.int32), tr["terminated"])


@pytest.mark.parametrize("hidden,n,dense,frame_skip,skip", [(64, 1000, True, 1, False), (32, 64, False, 1, False),
                                                            (128, 333, True, 3, False), (64, 16384, True, 1, False),
                                                            (64, 700, True, 1, True), (32, 333, True, 2, True)])
def test_horizon_kernel_equals_per_step_path(hidden, n, dense, frame_skip, skip):
    """fg_rollout_mlp (one launch per horizon, state in registers) against fg_policy_mlp_sample + fg_step per step:
    every rollout buffer, the final battle state and the episode statistics are bit-identical, over several horizons
    (ragged batch sizes: 1000 and 333 are not multiples of the 64 battles a CTA owns); skip = the fused
    FootsiesFrameSkipped stepping (fg_config.skip_unactionable) inside both."""
    _horizon_equals_per_step(hidden, n, dense, frame_skip, skip)


@pytest.mark.parametrize("hidden,n,dense,frame_skip,skip", [(64, 1000, True, 1, False), (32, 333, True, 2, True),
                                                            (64, 97, False, 1, False)])
def test_horizon_kernel_with_32_battle_warps_on_small_ragged_batches(monkeypatch, hidden, n, dense, frame_skip, skip):
    """From 131 072 battles up the whole-horizon kernel gives a warp two 16-row M tiles (32 battles, every lane simulates
    one).  That shape forced onto small ragged batches (FOOTSIES_B200_ROLLOUT_MT, read per launch) must not change a bit."""
    monkeypatch.setenv("FOOTSIES_B200_ROLLOUT_MT", "2")
    _horizon_equals_per_step(hidden, n, dense, frame_skip, skip)


def test_horizon_kernel_at_the_size_that_selects_32_battle_warps():
    """131 072 + 37 battles: the launcher picks the 32-battle-warp shape by itself, with a ragged last warp."""
    _horizon_equals_per_step(64, 131072 + 37, True, 1, False, horizon=24, rounds=2, min_launch_ratio=10)


def _horizon_equals_per_step(hidden, n, dense, frame_skip, skip, horizon=150, rounds=3, min_launch_ratio=50):
    from footsies_gym_b200 import FootsiesEnv
    from footsies_gym_b200.rollout import MLPPolicy, RolloutCollector
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible")
    dev = torch.device("cuda:0")
    torch.manual_seed(hidden + n)
    policy = MLPPolicy(hidden).to(dev)
    with torch.no_grad():
        for prm in policy.net.parameters():
            prm.mul_(2.0)
    # default horizon 150 > 120 frames: the statistics byte lanes are folded mid-horizon
    envs, cols = [], []
    for mode in ("step", "horizon"):
        env = FootsiesEnv(num_envs=n, device=dev, seed=3, dense_reward=dense, frame_skip=frame_skip, skip_unactionable=skip)
        envs.append(env)
        cols.append(RolloutCollector(env, policy, horizon=horizon, use_cuda_graph=False, fused=mode, seed=17))
    assert [c.mode for c in cols] == ["step", "horizon"]
    for r in range(rounds):
        outs = [c.collect() for c in cols]
        torch.cuda.synchronize()
        for k in ("obs", "actions", "logp", "rewards", "dones", "last_obs"):
            assert torch.equal(outs[0][k], outs[1][k]), (r, k)
        assert outs[0]["dones"].any() or horizon < 100
        s0, s1 = envs[0].get_state(), envs[1].get_state()
        assert s0.tobytes() == s1.tobytes(), r
        assert envs[0].episode_stats() == envs[1].episode_stats(), r
        assert torch.equal(envs[0].info_frame, envs[1].info_frame) and torch.equal(envs[0].info_misc, envs[1].info_misc)
    launches = [e.launch_count() for e in envs]
    assert launches[1] < launches[0] // min_launch_ratio    # 1 launch per horizon instead of 1 per step (+ the policy's)


@pytest.mark.parametrize("hidden,n,mirror,shared", [(64, 1000, False, False), (64, 777, True, True), (32, 130, True, False),
                                                    (128, 200, False, False)])
def test_self_play_rollout_with_a_policy_for_p2(hidden, n, mirror, shared):
    _self_play_rollout(hidden, n, mirror, shared)


def test_self_play_rollout_with_32_battle_warps(monkeypatch):
    monkeypatch.setenv("FOOTSIES_B200_ROLLOUT_MT", "2")
    _self_play_rollout(64, 777, True, True)


def _self_play_rollout(hidden, n, mirror, shared):
    """P2 driven by a second MLP policy (optionally on the mirrored observation, optionally the same network): the
    whole-horizon kernel equals the per-step path (fg_policy_mlp_sample + fg_policy_mlp_sample_p2 + fg_step) bit for bit,
    P2's log-probabilities equal torch's for the (mirrored) observation, and the collected actions replay."""
    from footsies_gym_b200 import FootsiesEnv
    from footsies_gym_b200.rollout import MLPPolicy, RolloutCollector, mirror_obs, _MIRROR_ACTION
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible")
    dev = torch.device("cuda:0")
    torch.manual_seed(hidden + n)
    p1 = MLPPolicy(hidden).to(dev)
    p2 = p1 if shared else MLPPolicy(hidden).to(dev)
    with torch.no_grad():
        for prm in list(p1.net.parameters()) + ([] if shared else list(p2.net.parameters())):
            prm.mul_(2.0)
    horizon = 130
    envs, cols = [], []
    for mode in ("step", "horizon"):
        env = FootsiesEnv(num_envs=n, device=dev, seed=3, opponent="self_play")
        envs.append(env)
        cols.append(RolloutCollector(env, p1, horizon=horizon, use_cuda_graph=False, fused=mode, seed=17, opponent_policy=p2,
                                     mirror_opponent=mirror))
    assert [c.mode for c in cols] == ["step", "horizon"]
    keys = ("obs", "actions", "logp", "rewards", "dones", "last_obs", "actions_p2", "logp_p2")
    for r in range(2):
        before = envs[0].get_state()
        outs = [c.collect() for c in cols]
        torch.cuda.synchronize()
        for k in keys:
            assert torch.equal(outs[0][k], outs[1][k]), (r, k)
        assert envs[0].get_state().tobytes() == envs[1].get_state().tobytes(), r
        assert envs[0].episode_stats() == envs[1].episode_stats(), r
    out = outs[1]
    assert out["dones"].any() and len(torch.unique(out["actions_p2"])) > 4
    # P2's log-probability is torch's log-softmax of its policy on the observation it was shown, at the action it chose
    mir = torch.tensor(_MIRROR_ACTION, device=dev)
    with torch.no_grad():
        for t in (0, horizon // 2, horizon - 1):
            o = mirror_obs(out["obs"][t]) if mirror else out["obs"][t]
            ref = torch.log_softmax(p2(o), dim=-1)
            chosen = mir[out["actions_p2"][t].long()] if mirror else out["actions_p2"][t].long()   # un-mirror: the policy's own pick
            assert float((ref.gather(1, chosen.unsqueeze(1)).squeeze(1) - out["logp_p2"][t]).abs().max()) < 2e-5
    # replay of the last horizon's actions from the state it started in
    twin = FootsiesEnv(num_envs=n, device=dev, seed=3, opponent="self_play")
    twin.reset()
    twin.set_state(before)
    for t in range(horizon):
        twin.step(out["actions"][t], out["actions_p2"][t])
        assert torch.equal(twin.reward, out["rewards"][t]) and torch.equal(twin.terminated, out["dones"][t]), t
    assert torch.equal(twin.obs, out["last_obs"])


def test_self_play_rollout_with_torch_policies_is_replayable():
    """The torch-op path (any callable policy) with an opponent policy on the mirrored observation: the collected actions
    replay into a fresh self-play env."""
    from footsies_gym_b200 import FootsiesEnv
    from footsies_gym_b200.rollout import MLPPolicy, RolloutCollector
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible")
    dev = torch.device("cuda:0")
    torch.manual_seed(1)
    n, horizon = 300, 90
    pol = MLPPolicy(64).to(dev)
    env = FootsiesEnv(num_envs=n, device=dev, seed=3, opponent="self_play")
    col = RolloutCollector(env, pol, horizon=horizon, use_cuda_graph=False, fused=False, opponent_policy=pol, mirror_opponent=True)
    assert col.mode == "torch"
    before = env.get_state()
    out = col.collect()
    assert int(out["actions_p2"].max()) <= 7 and len(torch.unique(out["actions_p2"])) > 4
    twin = FootsiesEnv(num_envs=n, device=dev, seed=3, opponent="self_play")
    twin.reset()
    twin.set_state(before)
    for t in range(horizon):
        if t > 0:
            assert torch.equal(twin.obs, out["obs"][t]), t
        twin.step(out["actions"][t], out["actions_p2"][t])
        assert torch.equal(twin.reward, out["rewards"][t]) and torch.equal(twin.terminated, out["dones"][t]), t
    assert torch.equal(twin.obs, out["last_obs"])


def test_horizon_kernel_rejects_other_configurations():
    from footsies_gym_b200 import FootsiesEnv
    from footsies_gym_b200.rollout import MLPPolicy, RolloutCollector
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible")
    dev = torch.device("cuda:0")
    policy = MLPPolicy(64).to(dev)
    env = FootsiesEnv(num_envs=64, device=dev, autoreset=False)
    with pytest.raises(ValueError):
        RolloutCollector(env, policy, horizon=8, fused="horizon")
    assert RolloutCollector(env, policy, horizon=8).mode == "step"
    assert RolloutCollector(FootsiesEnv(num_envs=64, device=dev), policy, horizon=8).mode == "horizon"


@pytest.mark.parametrize("hidden,n", [(32, 10007), (64, 10007), (128, 10007), (64, 50001)])
def test_fused_policy_kernel_matches_torch(hidden, n):
    """fg_policy_mlp_sample against the torch module it replaces: the log-probability it reports for the action it drew
    equals torch's log_softmax there (tolerance 2e-5 absolute: __expf-based tanh / exp, FMA contraction), and the
    actions it draws follow the policy's distribution."""
    from footsies_gym_b200.rollout import MLPPolicy
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible")
    dev = torch.device("cuda:0")
    torch.manual_seed(hidden)
    pol = MLPPolicy(hidden).to(dev)
    with torch.no_grad():
        for prm in pol.net.parameters():
            prm.mul_(3.0)                                   # a peaky, non-uniform policy
    # n is ragged (not a multiple of the 64 battles a CTA handles per pass); 50 001 battles make every CTA of the
    # persistent grid run several passes over its shared-memory buffers
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    obs = torch.stack([torch.randint(0, 4, (n,), generator=g, device=dev).float(), torch.randint(0, 4, (n,), generator=g, device=dev).float(),
                       torch.randint(0, 15, (n,), generator=g, device=dev).float(), torch.randint(0, 15, (n,), generator=g, device=dev).float(),
                       torch.randint(0, 56, (n,), generator=g, device=dev).float(), torch.randint(0, 56, (n,), generator=g, device=dev).float(),
                       torch.rand(n, generator=g, device=dev) * 9.2 - 4.6, torch.rand(n, generator=g, device=dev) * 9.2 - 4.6], dim=1).contiguous()
    actions = torch.zeros(n, dtype=torch.uint8, device=dev)
    logp = torch.zeros(n, dtype=torch.float32, device=dev)
    copy = torch.zeros_like(obs)
    base = torch.zeros(1, dtype=torch.int64, device=dev)
    with torch.no_grad():
        ref = torch.log_softmax(pol(obs), dim=-1)
    pol.fused_sample(obs, actions, logp, seed=5, counter=0, counter_base=base, obs_copy=copy)
    torch.cuda.synchronize()
    assert torch.equal(copy, obs) and int(actions.max()) <= 7
    got = ref.gather(1, actions.long().unsqueeze(1)).squeeze(1)
    assert float((got - logp).abs().max()) < 2e-5
    # determinism and the device-side counter: same (seed, counter) -> same draw; counter + base is what matters
    a2 = torch.zeros_like(actions)
    pol.fused_sample(obs, a2, None, seed=5, counter=0, counter_base=base)
    assert torch.equal(a2, actions)
    base.fill_(3)
    a3 = torch.zeros_like(actions)
    pol.fused_sample(obs, a3, None, seed=5, counter=0, counter_base=base)
    a4 = torch.zeros_like(actions)
    pol.fused_sample(obs, a4, None, seed=5, counter=3, counter_base=None)
    assert torch.equal(a3, a4) and not torch.equal(a3, actions)
    # distribution: one observation repeated, 200 000 draws, compared with the softmax probabilities
    m = 200_000
    one = obs[:1].expand(m, 8).contiguous()
    acts = torch.zeros(m, dtype=torch.uint8, device=dev)
    pol.fused_sample(one, acts, None, seed=9, counter=7)
    freq = torch.bincount(acts.long(), minlength=8).double() / m
    prob = ref[0].exp().double()
    assert float((freq - prob).abs().max()) < 5e-3, (freq.tolist(), prob.tolist())
