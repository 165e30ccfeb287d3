"""Parity at BASELINE.json's full sizes, through size-independent properties (the oracle cannot follow a million
battles frame by frame in test time):

  * configs[3] 1 Mi envs, random P1 vs bot: results do not depend on how the envs are sharded (one batch vs two
    half batches with global env indices), conservation laws of the episode statistics hold, state invariants hold,
    and a random sample of the million battles is followed exactly by per-env oracles;
  * configs[2] 65 536 envs self-play, K = 4 fused: identical to four masked K = 1 steps with the action repeated;
  * configs[4] 16 384 envs x 128-step horizon with a torch policy: every transition the collector stored is
    reproduced by replaying its actions (tests/test_rollout.py covers the oracle side at a smaller size).
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

INT_FIELDS = ["action_id", "action_frame", "hitstun", "guard", "vital", "hit_count", "buffer_id", "reserve_id",
              "is_input_backward", "is_reserve_prox", "shake", "attack_run", "hist_left", "hist_right"]


def _cuda():
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible")
    return torch.device("cuda:0")


def _states_equal(a, b, where):
    for f in INT_FIELDS + ["pos_x", "velocity_x"]:
        assert np.array_equal(a["f"][f], b["f"][f]), f"{where}: f.{f}"
    for f in ("frame", "recorded_input", "done", "cum_reward_index", "actor_input", "rng_state", "bot_queue", "p1_bot_memory"):
        assert np.array_equal(a[f], b[f]), f"{where}: {f}"


def test_config_d_one_million_envs_sharding_invariance_and_conservation(oracle):
    from footsies_gym_b200 import FootsiesEnv
    dev = _cuda()
    n, steps, seed = 1 << 20, 400, 17
    whole = FootsiesEnv(num_envs=n, device=dev, seed=seed)
    halves = [FootsiesEnv(num_envs=n // 2, device=dev, seed=seed, first_env_index=k * (n // 2)) for k in (0, 1)]
    gen = torch.Generator(device=dev)
    gen.manual_seed(99)
    for e in [whole] + halves:
        e.reset()
    # the reference's own specification, over every battle, every step (tests/invariants.py): observation_space bounds
    # (|position| <= 4.6), an episode's dense rewards sum to +-1, a guard break needs an already-empty guard bar
    from invariants import ReferenceInvariants
    inv = ReferenceInvariants(n, dev, dense=True)
    inv.reset(whole.obs, whole.info_frame, whole.info_misc)
    sample = np.sort(np.random.default_rng(5).choice(n, size=1024, replace=False))
    orcs = [oracle.OracleBatch(1, p2_bot=True, seed=seed, first_env_index=int(i)) for i in sample]
    for o in orcs:
        o.reset()
    sample_t = torch.from_numpy(sample).to(dev)
    for t in range(steps):
        a = torch.randint(0, 8, (n,), generator=gen, device=dev, dtype=torch.uint8)
        whole.step(a)
        halves[0].step(a[: n // 2])
        halves[1].step(a[n // 2:])
        inv.update(whole.obs, whole.reward, whole.terminated, whole.info_frame, whole.info_misc, where=f"step {t}")
        a_s = a[sample_t].cpu().numpy()
        for k, o in enumerate(orcs):
            o.step(a_s[k:k + 1])
        if t % 100 == 99 or t == steps - 1:
            # (1) sharding invariance, bit for bit, outputs included
            for k in (0, 1):
                sl = slice(k * (n // 2), (k + 1) * (n // 2))
                assert torch.equal(whole.obs[sl], halves[k].obs), f"step {t}: obs of shard {k}"
                assert torch.equal(whole.reward[sl], halves[k].reward) and torch.equal(whole.terminated[sl], halves[k].terminated)
                assert torch.equal(whole.state[:, sl], halves[k].state), f"step {t}: state planes of shard {k}"
            # (2) the sampled battles against their oracles
            obs = whole.obs[sample_t].cpu().numpy()
            rew = whole.reward[sample_t].cpu().numpy()
            term = whole.terminated[sample_t].cpu().numpy()
            frame = whole.info_frame[sample_t].cpu().numpy()
            for k, o in enumerate(orcs):
                tr = o.trace[0]
                assert np.array_equal(obs[k], tr["obs"]) and rew[k] == tr["reward"], (t, int(sample[k]))
                assert int(term[k]) == int(tr["terminated"]) and int(frame[k]) == int(tr["info_frame"]), (t, int(sample[k]))
    torch.cuda.synchronize()
    # (3) conservation laws of the statistics, and shards sum to the whole
    st = whole.episode_stats()
    parts = [h.episode_stats() for h in halves]
    for key in st:
        assert st[key] == parts[0][key] + parts[1][key], key
    assert st["episodes"] == st["p1_wins"] + st["p2_wins"] + st["double_ko"]
    assert st["env_frames"] + st["resets"] - n == n * steps          # every env did one thing per step: a frame or a reset
    done_now = int(whole.get_state()["done"].sum())
    assert st["resets"] - n == st["episodes"] - done_now             # every finished battle was restarted, except those still waiting
    assert st["hits"] + st["blocks"] + st["guard_breaks"] > 0 and st["episodes"] > n // 4
    # (4) state invariants over the whole million
    s = whole.get_state()
    f = s["f"]
    assert f["guard"].min() >= 0 and f["guard"].max() <= 3 and set(np.unique(f["vital"])) <= {0, 1}
    assert f["hitstun"].min() >= 0 and f["hitstun"].max() <= 30 and f["attack_run"].max() <= 59
    assert np.all(np.abs(f["pos_x"]) <= np.float32(4.6))              # footsies.py:167 (DASH_BACKWARD's 0.8 pushbox at the wall)
    assert inv.episodes == st["episodes"] and inv.breaks > 0 and inv.blocks > 1000
    assert np.all((f["vital"].min(axis=1) == 0) == (s["done"] == 1))
    for e in [whole] + halves:
        e.close()


def test_config_c_65536_self_play_fused_k4_equals_four_single_frames(oracle):
    """configs[2] at full size: the fused K = 4 launch against (a) four masked K = 1 steps of the same library, every
    battle, and (b) per-battle CPU oracles (repeat = 4) on a 1 024-battle sample of the 65 536, every macro step --
    plus the reference-held invariants (tests/invariants.py) over all of them."""
    from footsies_gym_b200 import FootsiesEnv
    from invariants import ReferenceInvariants
    dev = _cuda()
    n, steps, k = 65536, 300, 4
    fused = FootsiesEnv(num_envs=n, device=dev, opponent="self_play", frame_skip=k, seed=0)
    single = FootsiesEnv(num_envs=n, device=dev, opponent="self_play", frame_skip=1, seed=0)
    fused.reset()
    single.reset()
    sample = np.sort(np.random.default_rng(6).choice(n, size=1024, replace=False))
    sample_t = torch.from_numpy(sample).to(dev)
    orc = oracle.OracleBatch(len(sample), p2_bot=False, seed=0, threads=8)
    orc.reset()
    inv = ReferenceInvariants(n, dev, dense=True)
    inv.reset(fused.obs, fused.info_frame, fused.info_misc)
    gen = torch.Generator(device=dev)
    gen.manual_seed(7)
    cur1 = torch.randint(0, 8, (n,), generator=gen, device=dev, dtype=torch.uint8)
    cur2 = torch.randint(0, 8, (n,), generator=gen, device=dev, dtype=torch.uint8)
    for t in range(steps):
        # sticky inputs so that dashes, charged specials and guard breaks occur
        ch1 = torch.rand(n, generator=gen, device=dev) < 0.3
        ch2 = torch.rand(n, generator=gen, device=dev) < 0.3
        cur1 = torch.where(ch1, torch.randint(0, 8, (n,), generator=gen, device=dev, dtype=torch.uint8), cur1)
        cur2 = torch.where(ch2, torch.randint(0, 8, (n,), generator=gen, device=dev, dtype=torch.uint8), cur2)
        fused.step(cur1, cur2)
        inv.update(fused.obs, fused.reward, fused.terminated, fused.info_frame, fused.info_misc, where=f"macro step {t}")
        # the same macro step as masked single frames: a finished env only resets; a running env stops at its KO
        was_done = single.terminated.clone()
        total = torch.zeros(n, dtype=torch.float64, device=dev)
        active = torch.ones(n, dtype=torch.bool, device=dev)
        for j in range(k):
            single.set_step_mask(active)
            single.step(cur1, cur2)
            total += torch.where(active, single.reward.double(), torch.zeros_like(total))
            active = active & ~single.terminated & ~was_done
        single.set_step_mask(None)
        assert torch.equal(fused.state, single.state), f"macro step {t}: state planes"
        assert torch.equal(fused.obs, single.obs) and torch.equal(fused.terminated, single.terminated), f"macro step {t}"
        assert torch.equal(fused.info_frame, single.info_frame), f"macro step {t}"
        # rewards: the fused kernel sums in float64 and rounds once; tolerance for the per-frame float32 roundings
        assert float((fused.reward.double() - total).abs().max()) <= 1e-6, f"macro step {t}: reward"
        # (b) the sample against the CPU oracle stepping the same macro step (same action for up to 4 frames, stop at KO)
        tr = orc.step(cur1[sample_t].cpu().numpy(), cur2[sample_t].cpu().numpy(), repeat=k)
        assert np.array_equal(fused.obs[sample_t].cpu().numpy(), tr["obs"]), f"macro step {t}: obs vs oracle"
        assert np.array_equal(fused.reward[sample_t].cpu().numpy(), tr["reward"]), f"macro step {t}: reward vs oracle"
        assert np.array_equal(fused.terminated[sample_t].cpu().numpy().astype(np.int32), tr["terminated"]), f"macro step {t}"
        assert np.array_equal(fused.info_frame[sample_t].cpu().numpy(), tr["info_frame"]), f"macro step {t}"
        if t % 50 == 49:
            from parity import compare_states
            compare_states(fused.get_state()[sample], tr, where=f"macro step {t}", check_rng=False)
    a, b = fused.episode_stats(), single.episode_stats()
    for key in ("episodes", "p1_wins", "p2_wins", "double_ko", "episode_frames", "guard_breaks", "hits", "blocks", "env_frames"):
        assert a[key] == b[key], key
    assert a["episodes"] > 1000 and a["guard_breaks"] > 0
    fused.close()
    single.close()


def test_config_e_rollout_16384_envs_128_steps_is_replayable():
    from footsies_gym_b200 import FootsiesEnv
    from footsies_gym_b200.rollout import MLPPolicy, RolloutCollector
    dev = _cuda()
    n, horizon = 16384, 128
    torch.manual_seed(1)
    env = FootsiesEnv(num_envs=n, device=dev, seed=3)
    col = RolloutCollector(env, MLPPolicy().to(dev), horizon=horizon, use_cuda_graph=True)   # fused policy kernel, zero-copy
    col.collect()
    before = env.get_state()
    out = col.collect()
    torch.cuda.synchronize()
    ref = FootsiesEnv(num_envs=n, device=dev, seed=3)
    ref.reset()
    ref.set_state(before)
    for t in range(horizon):
        if t > 0:
            assert torch.equal(out["obs"][t], ref.obs), t
        ref.step(out["actions"][t])
        assert torch.equal(out["rewards"][t], ref.reward) and torch.equal(out["dones"][t], ref.terminated), t
    assert torch.equal(out["last_obs"], ref.obs)
    assert int(out["dones"].sum()) > 0 and len(torch.unique(out["actions"])) == 8
    env.close()
    ref.close()
