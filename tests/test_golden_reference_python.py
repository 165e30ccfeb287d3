"""Golden vectors recorded from the reference's own FootsiesEnv (tests/golden/make_golden.py) replayed into
(a) the CPU oracle -- pins the oracle's restatement of footsies.py -- and (b) the CUDA path through the C ABI.

Every value must match exactly: integers bit for bit, observations as the float64 image of the fp32 value,
rewards as the reference's Python float (oracle: identical double; kernel: its fp32 cast).
"""
import glob
import os

import numpy as np
import pytest

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_python_*.npz")))


def _ids(paths):
    return [os.path.basename(p)[len("ref_python_"):-4] for p in paths]


def test_golden_files_exist():
    assert len(GOLDEN) >= 4


class OracleDriver:
    def __init__(self, ob, dense, delay, p2_remote):
        self.b = ob.OracleBatch(1, p2_bot=not p2_remote, dense_reward=dense, frame_delay=delay, autoreset=False, seed=0)

    def seed(self, v):
        self.b.seed(v)

    def reset(self):
        return self._out(self.b.reset()[0])

    def step(self, a1, a2):
        return self._out(self.b.step([a1], [a2])[0])

    @staticmethod
    def _out(t):
        return dict(obs=t["obs"].astype(np.float64), reward=float(t["reward_f64"]), reward32=np.float32(t["reward"]),
                    terminated=int(t["terminated"]), frame=int(t["info_frame"]),
                    action=t["info_action"].tolist(), hitstun=t["info_hitstun"].tolist())


class KernelDriver:
    def __init__(self, dense, delay, p2_remote):
        from footsies_gym_b200 import FootsiesEnv
        self.env = FootsiesEnv(num_envs=1, device="cuda:0", dense_reward=dense, frame_delay=delay, autoreset=False,
                               opponent="remote" if p2_remote else None, seed=0)
        self.p2_remote = p2_remote

    def seed(self, v):
        self.env.seed(v)

    def reset(self):
        obs, info = self.env.reset()
        return self._out(obs, None, None, info)

    def step(self, a1, a2):
        import torch
        obs, reward, term, trunc, info = self.env.step(torch.tensor([a1], dtype=torch.uint8),
                                                       torch.tensor([a2], dtype=torch.uint8) if self.p2_remote else None)
        assert not bool(trunc[0])
        return self._out(obs, reward, term, info)

    @staticmethod
    def _out(obs, reward, term, info):
        o = np.concatenate([obs[k][0].cpu().numpy() for k in ("guard", "move", "move_frame", "position")]).astype(np.float64)
        r32 = np.float32(0.0) if reward is None else np.float32(reward[0].item())
        return dict(obs=o, reward=None, reward32=r32, terminated=0 if term is None else int(term[0]),
                    frame=int(info["frame"][0]), action=[int(info["p1_action"][0]), int(info["p2_action"][0])],
                    hitstun=[int(info["p1_hitstun"][0]), int(info["p2_hitstun"][0])])


def replay(path, driver):
    g = np.load(path)
    ops = g["ops"]
    checked = 0
    for j, (kind, a, b, ei) in enumerate(ops.tolist()):
        if kind == 2:
            driver.seed(a)
            continue
        out = driver.reset() if kind == 1 else driver.step(a, b)
        if ei < 0:
            continue
        where = f"{os.path.basename(path)} op {j} (expectation {ei})"
        assert g["exp_kind"][ei] == kind, where
        assert np.array_equal(out["obs"], g["exp_obs"][ei]), (where, out["obs"], g["exp_obs"][ei])
        assert out["frame"] == g["exp_frame"][ei], where
        assert out["action"] == g["exp_action"][ei].tolist(), where
        assert out["hitstun"] == g["exp_hitstun"][ei].tolist(), where
        if kind == 0:
            if out["reward"] is not None:
                assert out["reward"] == float(g["exp_reward"][ei]), (where, out["reward"], g["exp_reward"][ei])
            assert out["reward32"] == np.float32(g["exp_reward"][ei]), (where, out["reward32"], g["exp_reward"][ei])
            assert out["terminated"] == g["exp_terminated"][ei], where
            assert g["exp_truncated"][ei] == 0
        checked += 1
    assert checked == len(g["exp_kind"]), (checked, len(g["exp_kind"]))
    return checked


@pytest.mark.parametrize("path", GOLDEN, ids=_ids(GOLDEN))
def test_oracle_matches_reference_python(oracle, path):
    dense, delay, p2_remote, _ = np.load(path)["config"].tolist()
    n = replay(path, OracleDriver(oracle, bool(dense), int(delay), bool(p2_remote)))
    assert n > 1000


@pytest.mark.parametrize("path", GOLDEN, ids=_ids(GOLDEN))
def test_transliterated_reference_engine_matches_reference_python(path):
    """The same goldens with the reference's OWN C# battle code (oracle/_ref, tools/cs2cpp.py) playing the game: what the
    unmodified reference FootsiesEnv computed in Python is reproduced when its C# engine -- not the hand-written oracle --
    produces the states (closes the triangle reference Python <-> reference C# <-> oracle on these tapes)."""
    import types
    import ref_binding
    if not ref_binding.available():
        pytest.skip("oracle/_ref not built")
    dense, delay, p2_remote, _ = np.load(path)["config"].tolist()
    n = replay(path, OracleDriver(types.SimpleNamespace(OracleBatch=ref_binding.RefBatch), bool(dense), int(delay), bool(p2_remote)))
    assert n > 1000


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=_ids(GOLDEN))
def test_kernel_matches_reference_python(path):
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible")
    dense, delay, p2_remote, _ = np.load(path)["config"].tolist()
    n = replay(path, KernelDriver(bool(dense), int(delay), bool(p2_remote)))
    assert n > 1000
