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


@pytest.mark.parametrize("k,p2_bot", pc.FUSED_PARAMS)
def test_fused_frame_skip(oracle, k, p2_bot):
    pc.case_fused_frame_skip(make_env, oracle, k, p2_bot)


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


@pytest.mark.parametrize("min_envs", ["0", "1000000000"], ids=["large_shapes_on_small_batches", "small_shape_only"])
def test_cta_shapes_are_interchangeable(min_envs):
    """The step kernel picks a CTA shape by batch size (768 / 1024-thread pipeline groups from 0.75 Mi envs, 256-thread
    CTAs below).  Forcing each choice onto ragged, small batches (few chunks per pipeline group, tail chunks, masked
    steps, fused K) must not change a single bit: tools/sanitize_check.py compares against the oracle."""
    import os
    import subprocess
    import sys
    _require_cuda()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, FOOTSIES_B200_LARGE_SHAPE_MIN_ENVS=min_envs, SAN_STEPS="120")
    res = subprocess.run([sys.executable, os.path.join(root, "tools", "sanitize_check.py")], env=env, capture_output=True,
                         text=True, timeout=600)
    assert res.returncode == 0 and "sanitize_check done" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]
