"""The batched wrappers against golden vectors recorded from the reference's own wrapper classes
(footsies_gym/wrappers/*.py, run unmodified by tests/golden/make_golden.py).

Integer-valued fields and termination must match exactly; normalised floats and accumulated rewards are float64
in the reference and float32 here: tolerance 1e-6 absolute (stated)."""
import glob
import os

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = sorted(glob.glob(os.path.join(HERE, "golden", "ref_wrappers_*.npz")))
TOL = 1e-6


def build_chain(base, chain):
    from footsies_gym_b200.wrappers import (FootsiesActionCombinationsDiscretized, FootsiesFrameSkipped,
                                            FootsiesNormalized, FootsiesStatistics)
    env, stats = base, None
    for w in chain:
        if w == "normalized":
            env = FootsiesNormalized(env)
        elif w == "frame_skipped":
            env = FootsiesFrameSkipped(env)
        elif w == "discretized":
            env = FootsiesActionCombinationsDiscretized(env)
        elif w == "statistics":
            env = stats = FootsiesStatistics(env)
    return env, stats


def replay(path, base):
    g = np.load(path)
    chain = [str(c) for c in g["chain"]]
    env, stats = build_chain(base, chain)
    n_checked = 0
    for j, (kind, a) in enumerate(g["calls"].tolist()):
        where = f"{os.path.basename(path)} call {j}"
        if kind == 1:
            obs, info = env.reset(seed=None, options=None)
            reward, terminated = 0.0, 0
        else:
            act = torch.tensor([a], dtype=torch.uint8) if "discretized" not in chain else torch.tensor([a])
            obs, reward, terminated, truncated, info = env.step(act)
            reward, terminated = float(reward[0]), int(terminated[0])
        exp = g["exp_obs"][j]
        got = [float(obs["guard"][0, 0]), float(obs["guard"][0, 1]), float(obs["move"][0, 0]), float(obs["move"][0, 1])]
        mf = obs["move_frame"]
        got += [float(mf[0, 0]), float(mf[0, 1]) if mf.shape[1] > 1 else float("nan")]
        got += [float(obs["position"][0, 0]), float(obs["position"][0, 1])]
        assert got[2:4] == exp[2:4].tolist(), where                      # move indices: exact
        assert np.allclose(np.array(got), exp, atol=TOL, rtol=0, equal_nan=True), (where, got, exp.tolist())
        assert abs(reward - float(g["exp_reward"][j])) <= TOL, (where, reward, g["exp_reward"][j])
        assert terminated == int(g["exp_terminated"][j]), where
        assert int(info["frame"][0]) == int(g["exp_frame"][j]), where
        n_checked += 1
    if stats is not None:
        assert stats.metric_special_moves_per_episode == g["special_moves_per_episode"].tolist()
        assert stats.metric_special_moves_from_neutral_per_episode == g["special_moves_from_neutral_per_episode"].tolist()
    return n_checked


def test_wrapper_goldens_exist():
    assert len(GOLDEN) >= 3


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[13:-4] for p in GOLDEN])
def test_wrappers_on_cpu_stand_in(oracle, path):
    from oracle_env import OracleTorchEnv
    assert replay(path, OracleTorchEnv(num_envs=1, autoreset=False, seed=0)) > 1000


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[13:-4] for p in GOLDEN])
def test_wrappers_on_host_compiled_kernel_logic(path):
    """The same goldens through the device frame logic compiled for the host: FootsiesFrameSkipped takes its fused path
    there (the skip loop of step_kernel, mirrored by tests/host_emulation), so the reference wrapper's own outputs pin it."""
    from kernel_host import HostKernelTorchEnv
    from footsies_gym_b200.wrappers import FootsiesFrameSkipped
    base = HostKernelTorchEnv(num_envs=1, autoreset=False, seed=0)
    assert FootsiesFrameSkipped(base).fused
    base.set_skip_unactionable(False)
    assert replay(path, base) > 1000


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[13:-4] for p in GOLDEN])
def test_wrappers_on_gpu(path):
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible")
    from footsies_gym_b200 import FootsiesEnv
    assert replay(path, FootsiesEnv(num_envs=1, device="cuda:0", autoreset=False, seed=0)) > 1000


@pytest.mark.gpu
def test_step_mask_freezes_unselected_envs(oracle):
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible")
    from footsies_gym_b200 import FootsiesEnv
    from parity import compare_state_and_outputs
    n = 600
    rng = np.random.default_rng(9)
    env = FootsiesEnv(num_envs=n, device="cuda:0", seed=2)
    orc = oracle.OracleBatch(n, p2_bot=True, seed=2)
    env.reset()
    orc.reset()
    for t in range(60):
        a = rng.integers(0, 8, size=n, dtype=np.uint8)
        env.step(torch.from_numpy(a))
        orc.step(a)
    before = env.get_state().copy()
    obs_before = env.obs.clone()
    mask = torch.from_numpy(rng.random(n) < 0.5)
    env.set_step_mask(mask)
    a = rng.integers(0, 8, size=n, dtype=np.uint8)
    env.step(torch.from_numpy(a))
    env.set_step_mask(None)
    after = env.get_state()
    m = mask.numpy()
    assert np.array_equal(after[~m], before[~m])                       # untouched
    assert torch.equal(env.obs[~mask].cpu(), obs_before[~mask].cpu())
    orc.step(a)                                                        # oracle steps everybody: compare the stepped half
    st = after.copy()
    from parity import compare_states
    compare_states(st[m], orc.trace[m], where="masked step")
    env.close()


def test_frame_skipped_batch_matches_single_env_semantics(oracle):
    """_is_obs_skippable on a batch == the reference predicate applied row by row (frame_skip.py:56-66)."""
    from footsies_gym_b200.moves import FOOTSIES_MOVE_INDEX_TO_MOVE, FootsiesMove
    from footsies_gym_b200.wrappers import FootsiesFrameSkipped
    from oracle_env import OracleTorchEnv
    w = FootsiesFrameSkipped(OracleTorchEnv(num_envs=1))
    hit_guard = {FootsiesMove.DAMAGE, FootsiesMove.GUARD_STAND, FootsiesMove.GUARD_CROUCH, FootsiesMove.GUARD_M,
                 FootsiesMove.GUARD_BREAK}
    rng = np.random.default_rng(0)
    moves = rng.integers(0, 15, size=(500, 2))
    mf = rng.integers(0, 3, size=(500, 2)).astype(np.float32)
    obs = {"move": torch.from_numpy(moves.astype(np.float32)), "move_frame": torch.from_numpy(mf)}
    got = w._is_obs_skippable(obs).tolist()
    exp = [bool((mf[i, 0] != 0.0 and FOOTSIES_MOVE_INDEX_TO_MOVE[moves[i, 1]] not in hit_guard)
                or FOOTSIES_MOVE_INDEX_TO_MOVE[moves[i, 0]] == FootsiesMove.DAMAGE) for i in range(500)]
    assert got == exp


def test_get_dict_obs_from_vector_obs_round_trip():
    """footsies_gym/utils.py:7-40 on batches: normalised + flattened observations come back as the original dictionary;
    the flattened layout is gymnasium's (one-hot MultiDiscrete entries, Dict entries in declaration order)."""
    import torch
    from footsies_gym_b200.spaces import Box, Dict, footsies_observation_space
    from footsies_gym_b200.utils import flatdim, flatten_observation, get_dict_obs_from_vector_obs
    from footsies_gym_b200.wrappers import FOOTSIES_MOVE_INDEX_TO_MOVE
    g = torch.Generator().manual_seed(0)
    n = 257
    move = torch.randint(0, 15, (n, 2), generator=g).float()
    dur = torch.tensor([float(m.value.duration) for m in FOOTSIES_MOVE_INDEX_TO_MOVE])[move.long()]
    obs = {"guard": torch.randint(0, 4, (n, 2), generator=g).float(), "move": move,
           "move_frame": torch.floor(torch.rand((n, 2), generator=g) * dur),
           "position": (torch.rand((n, 2), generator=g) * 9.2 - 4.6)}
    space = footsies_observation_space()
    assert flatdim(space) == 4 + 4 + 15 + 15 + 2 + 2
    # (1) plain flatten -> unflatten
    flat = flatten_observation(space, obs)
    assert flat.shape == (n, 42) and bool((flat[:, :8].sum(dim=1) == 2).all())
    back = get_dict_obs_from_vector_obs(flat, flattened=True, unflattenend_observation_space=space, normalized=False)
    for k in obs:
        assert torch.equal(back[k], obs[k]), k
    # (2) normalised (guard too) and flattened, as a FootsiesNormalized + FlattenObservation stack would deliver it
    norm = {"guard": obs["guard"] / 3.0, "move": obs["move"], "move_frame": obs["move_frame"] / dur, "position": obs["position"] / 4.6}
    nspace = Dict({"guard": Box(0.0, 1.0, (2,)), "move": space["move"], "move_frame": Box(0.0, 1.0, (2,)), "position": Box(-1.0, 1.0, (2,))})
    back = get_dict_obs_from_vector_obs(flatten_observation(nspace, norm), True, nspace, normalized=True, normalized_guard=True)
    for k in obs:
        assert torch.allclose(back[k], obs[k], atol=1e-5), k
    # (3) a single unbatched observation, dictionary form, guard not normalised
    one = {k: v[3] for k, v in norm.items()}
    one["guard"] = obs["guard"][3]
    back = get_dict_obs_from_vector_obs(one, flattened=False, normalized=True, normalized_guard=False)
    for k in obs:
        assert torch.allclose(back[k], obs[k][3], atol=1e-5), k
    import pytest
    with pytest.raises(ValueError):
        get_dict_obs_from_vector_obs(flat, flattened=True)
    with pytest.raises(ValueError):
        get_dict_obs_from_vector_obs(flat, flattened=False)
