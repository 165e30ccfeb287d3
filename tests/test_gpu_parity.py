"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on the same seeded
action tapes -- every env, every frame, every field (see tests/parity.py for the bar)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

pytestmark = pytest.mark.gpu


def _require_cuda():
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible")


import parity_cases as pc
from parity_cases import tape_sticky, tape_uniform  # noqa: F401


def make_env(**kw):
    from footsies_gym_b200 import FootsiesEnv
    _require_cuda()
    return FootsiesEnv(device="cuda:0", **kw)


CASES = [name for name in dir(pc) if name.startswith("case_") and name != "case_fused_frame_skip"]


@pytest.mark.parametrize("name", CASES, ids=[c[5:] for c in CASES])
def test_parity_case(oracle, name):
    getattr(pc, name)(make_env, oracle)


@pytest.mark.parametrize("name", ["case_config_b_4096_envs_random_vs_bot_every_frame", "case_both_bots_by_example",
                                  "case_input_personalities_self_play"])
def test_parity_case_directly_against_the_transliterated_reference(name):
    """The CUDA path against the reference's OWN battle code (oracle/_ref: Assets/Script/*.cs transliterated by
    tools/cs2cpp.py; the prebuilt library travels to the GPU box) without the hand-written oracle in between: the trace
    gate of BASELINE configs[1], by_example and the self-play personalities, every battle, every frame, every field."""
    import types
    import ref_binding
    if not ref_binding.available():
        pytest.skip("oracle/_ref not built")
    _require_cuda()
    ref = types.SimpleNamespace(OracleBatch=ref_binding.RefBatch)
    getattr(pc, name)(make_env, ref, scale=0.125)


@pytest.mark.parametrize("k,p2_bot", pc.FUSED_PARAMS)
def test_fused_frame_skip(oracle, k, p2_bot):
    pc.case_fused_frame_skip(make_env, oracle, k, p2_bot)


@pytest.mark.parametrize("by_example", [False, True])
def test_masked_hard_reset_and_reseed(oracle, by_example):
    """RESET in mid-round and SEED on a subset of the battles.  by_example: P1's spectator-wrapped bot is never Reset()
    (BattleCore.cs:274) -- it keeps its queues and decides on the state it recorded last (found by tests/test_oracle_vs_ref.py)."""
    from footsies_gym_b200 import FootsiesEnv
    from parity import compare_state_and_outputs
    _require_cuda()
    rng = np.random.default_rng(8)
    n = 777                                  # ragged: not a multiple of the warp or CTA size
    env = FootsiesEnv(num_envs=n, device="cuda:0", seed=1, by_example=by_example)
    orc = oracle.OracleBatch(n, p1_bot=by_example, p2_bot=True, seed=1)
    env.reset()
    orc.reset()
    for t in range(400):
        a = rng.integers(0, 8, size=n, dtype=np.uint8)
        env.step(None if by_example else torch.from_numpy(a))
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


@pytest.mark.parametrize("frame_skip,p2_bot,n", [(1, True, 3000), (3, True, 1000), (1, False, 1000)])
def test_frame_skipped_fused_into_the_step_kernel(frame_skip, p2_bot, n):
    """fg_config.skip_unactionable against the FootsiesFrameSkipped loop of masked steps (wrappers/frame_skip.py:68-80)."""
    pc.frame_skipped_fused_vs_masked_loop(make_env, frame_skip, p2_bot, n=n, steps=300)


@pytest.mark.parametrize("chunk_envs", [None, 256])
def test_host_buffer_step_matches_oracle(oracle, monkeypatch, chunk_envs):
    """FootsiesEnv.step_host / reset_host (fg_step_host_compact: compact 27-byte host layout) against the oracle; with
    chunk_envs = 256 the 1000 battles go through the sliced two-stream pipeline (3 whole slices + a ragged one)."""
    from footsies_gym_b200 import FootsiesEnv
    _require_cuda()
    if chunk_envs:
        monkeypatch.setenv("FOOTSIES_B200_HOST_CHUNK_ENVS", str(chunk_envs))
    n = 1000
    rng = np.random.default_rng(6)
    env = FootsiesEnv(num_envs=n, seed=5)
    orc = oracle.OracleBatch(n, p2_bot=True, seed=5)
    obs, info = env.reset_host()
    tr = orc.reset()
    keys = ("guard", "move", "move_frame", "position")
    assert [obs[k].dtype for k in keys] == [torch.uint8] * 3 + [torch.float32]
    assert np.array_equal(np.concatenate([obs[k].numpy() for k in keys], 1), tr["obs"])
    assert np.array_equal(info["frame"].numpy(), tr["info_frame"])
    for t in range(200):
        a = rng.integers(0, 8, size=n, dtype=np.uint8)
        if t % 2:
            a = torch.from_numpy(a).pin_memory()      # pinned actions are uploaded from where they are
        obs, reward, term, trunc, info = env.step_host(a)
        tr = orc.step(np.asarray(a))
        assert np.array_equal(np.concatenate([obs[k].numpy() for k in keys], 1), tr["obs"])
        assert np.array_equal(reward.numpy(), tr["reward"])
        assert np.array_equal(term.numpy().astype(np.int32), tr["terminated"])
        assert not trunc.any()
        assert np.array_equal(info["frame"].numpy(), tr["info_frame"])
        assert np.array_equal(np.stack([info[k].numpy() for k in ("p1_action", "p2_action", "p1_hitstun", "p2_hitstun")], 1),
                              np.concatenate([tr["info_action"], tr["info_hitstun"]], 1))
        # the device-resident tensors hold the same step
        assert np.array_equal(env.obs.cpu().numpy(), tr["obs"])
    env.close()


def test_host_buffer_f32_layout_and_self_play_slices(oracle, monkeypatch):
    """fg_step_host (device layout: obs f32 [N][8]) through the sliced pipeline, self-play with an odd batch size."""
    import ctypes as C
    from footsies_gym_b200 import FootsiesEnv, _capi
    _require_cuda()
    monkeypatch.setenv("FOOTSIES_B200_HOST_CHUNK_ENVS", "512")
    n = 1337
    rng = np.random.default_rng(9)
    env = FootsiesEnv(num_envs=n, opponent="self_play", seed=1)
    orc = oracle.OracleBatch(n, p2_bot=False, seed=1)
    env.reset()
    orc.reset()
    obs = np.zeros((n, 8), np.float32); reward = np.zeros(n, np.float32); term = np.zeros(n, np.uint8)
    frame = np.zeros(n, np.int32); misc = np.zeros((n, 4), np.uint8)
    ptr = lambda x: C.c_void_p(x.ctypes.data)
    for t in range(150):
        a1 = rng.integers(0, 8, size=n, dtype=np.uint8)
        a2 = rng.integers(0, 8, size=n, dtype=np.uint8)
        _capi.check(env._lib.fg_step_host(env._handle, ptr(a1), ptr(a2), ptr(obs), ptr(reward), ptr(term), ptr(frame),
                                          ptr(misc), None))
        tr = orc.step(a1, a2)
        assert np.array_equal(obs, tr["obs"]) and np.array_equal(reward, tr["reward"])
        assert np.array_equal(term.astype(np.int32), tr["terminated"]) and np.array_equal(frame, tr["info_frame"])
        assert np.array_equal(misc, np.concatenate([tr["info_action"], tr["info_hitstun"]], 1))
        # and the compact layout of the same state
        o2, r2, t2, _, i2 = env.step_host(a1, a2)
        tr = orc.step(a1, a2)
        assert np.array_equal(np.concatenate([o2[k].numpy() for k in ("guard", "move", "move_frame", "position")], 1), tr["obs"])
        assert np.array_equal(r2.numpy(), tr["reward"])
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


@pytest.mark.parametrize("min_envs,pdl_min_envs", [("0", "0"), ("1000000000", "0"), ("0", "1000000000")],
                         ids=["large_shapes_on_small_batches_with_pdl", "small_shape_only_with_pdl", "large_shapes_without_pdl"])
def test_cta_shapes_are_interchangeable(min_envs, pdl_min_envs):
    """The step kernel picks a CTA shape by batch size (768 / 1024-thread pipeline groups from 0.75 Mi envs, 256-thread
    CTAs below).  Forcing each choice onto ragged, small batches (few chunks per pipeline group, tail chunks, masked
    steps, fused K) must not change a single bit: tools/sanitize_check.py compares against the oracle."""
    import os
    import subprocess
    import sys
    _require_cuda()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    # ... and with / without programmatic dependent launch (by default the fused-K kernels only use it from 262 144 battles)
    env = dict(os.environ, FOOTSIES_B200_LARGE_SHAPE_MIN_ENVS=min_envs, FOOTSIES_B200_PDL_MIN_ENVS=pdl_min_envs, SAN_STEPS="120")
    res = subprocess.run([sys.executable, os.path.join(root, "tools", "sanitize_check.py")], env=env, capture_output=True,
                         text=True, timeout=600)
    assert res.returncode == 0 and "sanitize_check done" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]
