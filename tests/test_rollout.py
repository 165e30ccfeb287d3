"""BASELINE.json configs[4]: PPO-style rollout with a torch MLP policy reading the observation tensor in place.
The collected transitions must be exactly what a fresh env produces when the recorded actions are replayed into it
(and into the CPU oracle), both for the eager loop and for the CUDA-graph-captured horizon."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("use_graph", [False, True])
def test_rollout_is_replayable(oracle, use_graph):
    from footsies_gym_b200 import FootsiesEnv
    from footsies_gym_b200.rollout import MLPPolicy, RolloutCollector
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible")
    n, horizon, rounds = 512, 64, 3
    torch.manual_seed(0)
    dev = torch.device("cuda:0")
    env = FootsiesEnv(num_envs=n, device=dev, seed=11)
    policy = MLPPolicy().to(dev)
    col = RolloutCollector(env, policy, horizon=horizon, use_cuda_graph=use_graph)
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
    col2 = RolloutCollector(env2, policy, horizon=horizon, use_cuda_graph=False)
    out = col2.collect()
    orc.reset()
    for t in range(horizon):
        tr = orc.step(out["actions"][t].cpu().numpy())
        assert np.array_equal(out["rewards"][t].cpu().numpy(), tr["reward"])
        assert np.array_equal(out["dones"][t].cpu().numpy().astype(np.int32), tr["terminated"])
