"""CPU-only tests: host logic of the front-end, the C-ABI library loads and exports every symbol the header
declares, it refuses to run without a GPU, and the multi-GPU plumbing works over gloo with world_size 2."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_are_all_exported():
    from footsies_gym_b200 import _capi
    header = open(os.path.join(ROOT, "include", "footsies_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(fg_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_capi.EXPORTS), declared ^ set(_capi.EXPORTS)
    lib = _capi.load()
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.fg_abi_version() == 2


def test_struct_layouts_match_header_sizes():
    from footsies_gym_b200 import _capi
    assert C.sizeof(_capi.FgConfig) == 48
    assert C.sizeof(_capi.FgBuffers) == 8 + 8 * (4 + 1 + 2 + 5 + 1)
    assert C.sizeof(_capi.FgFighterState) == 72
    assert C.sizeof(_capi.FgEnvState) == 2 * 72 + 4 * 14
    assert _capi.env_state_dtype().itemsize == C.sizeof(_capi.FgEnvState)


def test_ctypes_structs_match_the_header_as_gcc_sees_it(tmp_path):
    """Every struct of include/footsies_b200.h, compiled as plain C: size and every field offset equal the ctypes mirror
    in footsies_gym_b200/_capi.py (the header is the contract, the binding must not drift)."""
    from footsies_gym_b200 import _capi
    pairs = {"fg_config": _capi.FgConfig, "fg_buffers": _capi.FgBuffers, "fg_fighter_state": _capi.FgFighterState,
             "fg_env_state": _capi.FgEnvState, "fg_host_outputs": _capi.FgHostOutputs,
             "fg_rollout_buffers": _capi.FgRolloutBuffers}
    lines = ["#include <stdio.h>", "#include <stddef.h>", '#include "footsies_b200.h"', "int main(void) {"]
    for cname, cls in pairs.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ["return 0; }"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)],
                   check=True)
    got = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for cname, cls in pairs.items():
        assert int(got[cname]) == C.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(cls, fname).offset, (cname, fname)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    from footsies_gym_b200 import FootsiesEnv, _capi
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        FootsiesEnv(num_envs=4)
    cfg = _capi.FgConfig(struct_size=C.sizeof(_capi.FgConfig), num_envs=4, device=0, p2_bot=1, dense_reward=1,
                         frame_skip=1, autoreset=1, stale_intro_input=1)
    h = C.c_void_p()
    rc = _capi.load().fg_create(C.byref(cfg), C.byref(h))
    assert rc == -4 and b"no CPU fallback" in _capi.load().fg_last_error()


def test_argument_validation_without_gpu():
    from footsies_gym_b200 import _capi
    lib = _capi.load()
    h = C.c_void_p()
    bad = _capi.FgConfig(struct_size=4, num_envs=4)
    assert lib.fg_create(C.byref(bad), C.byref(h)) == -1
    bad = _capi.FgConfig(struct_size=C.sizeof(_capi.FgConfig), num_envs=0, frame_skip=1)
    assert lib.fg_create(C.byref(bad), C.byref(h)) == -1
    bad = _capi.FgConfig(struct_size=C.sizeof(_capi.FgConfig), num_envs=4, frame_skip=0)
    assert lib.fg_create(C.byref(bad), C.byref(h)) == -1
    assert lib.fg_step(None, None) == -1


def test_algorithmic_bytes():
    from footsies_gym_b200 import _capi
    lib = _capi.load()

    def nbytes(p1, p2):
        cfg = _capi.FgConfig(struct_size=C.sizeof(_capi.FgConfig), num_envs=1, p1_bot=p1, p2_bot=p2, frame_skip=1)
        return lib.fg_algorithmic_bytes_per_env_step(C.byref(cfg))
    # state planes read + written (3 without bots, 4 with the RNG plane) + actions + obs 32 + reward 4 + term 1 + info 8
    assert nbytes(0, 0) == 96 + 2 + 45
    assert nbytes(0, 1) == 128 + 1 + 45
    assert nbytes(1, 1) == 128 + 0 + 45


def test_action_conversion():
    from footsies_gym_b200.env import _as_bitmask
    assert _as_bitmask((True, False, True), 1, "cpu").tolist() == [5]
    assert _as_bitmask((0, 1, 1), 1, "cpu").tolist() == [6]
    assert _as_bitmask(np.array([[1, 0, 0], [0, 1, 0], [1, 1, 1]]), 3, "cpu").tolist() == [1, 2, 7]
    assert _as_bitmask(torch.tensor([0, 7, 3], dtype=torch.int64), 3, "cpu").tolist() == [0, 7, 3]
    with pytest.raises(ValueError):
        _as_bitmask(np.zeros((2, 3)), 3, "cpu")
    with pytest.raises(ValueError):
        _as_bitmask(np.zeros(4), 3, "cpu")


def test_spaces_match_reference_constants():
    # footsies.py:153-174
    from footsies_gym_b200.moves import FootsiesMove
    from footsies_gym_b200.spaces import footsies_action_space, footsies_observation_space
    relevant = [m for m in FootsiesMove if m.name not in ("WIN", "DEAD")]
    sp = footsies_observation_space(len(relevant), max(m.value.duration for m in relevant))
    assert list(sp.keys()) == ["guard", "move", "move_frame", "position"]
    assert sp["guard"].nvec.tolist() == [4, 4] and sp["move"].nvec.tolist() == [15, 15]
    assert float(sp["move_frame"].high[0]) == 55.0 and float(sp["move_frame"].low[0]) == 0.0
    assert abs(float(sp["position"].high[0]) - 4.6) < 1e-6 and sp["position"].shape == (2,)
    assert footsies_action_space().n == 3


def test_moves_table_matches_reference_values():
    # footsies_gym/moves.py:13-29, transcribed here as the known answers
    from footsies_gym_b200.moves import FOOTSIES_MOVE_ID_TO_INDEX, FootsiesMove
    exp = {"STAND": (0, 24, 0, 0, 0), "FORWARD": (1, 24, 0, 0, 0), "BACKWARD": (2, 24, 0, 0, 0),
           "DASH_FORWARD": (10, 16, 0, 0, 0), "DASH_BACKWARD": (11, 22, 0, 0, 0),
           "N_ATTACK": (100, 22, 4, 2, 16), "B_ATTACK": (105, 21, 3, 3, 15), "N_SPECIAL": (110, 44, 11, 4, 29),
           "B_SPECIAL": (115, 55, 2, 6, 47), "DAMAGE": (200, 17, 0, 0, 0), "GUARD_M": (301, 23, 0, 0, 0),
           "GUARD_STAND": (305, 15, 0, 0, 0), "GUARD_CROUCH": (306, 15, 0, 0, 0), "GUARD_BREAK": (310, 36, 0, 0, 0),
           "GUARD_PROXIMITY": (350, 1, 0, 0, 0), "DEAD": (500, 500, 0, 0, 0), "WIN": (510, 33, 0, 0, 0)}
    assert [m.name for m in FootsiesMove] == list(exp)
    for m in FootsiesMove:
        v = m.value
        assert (v.id, v.duration, v.startup, v.active, v.recovery) == exp[m.name], m
    assert FOOTSIES_MOVE_ID_TO_INDEX[350] == 14 and FOOTSIES_MOVE_ID_TO_INDEX[110] == 7
    assert FootsiesMove.N_ATTACK.in_active(4) and FootsiesMove.N_ATTACK.in_startup(3) and FootsiesMove.N_ATTACK.in_recovery(6)


def test_shard_ranges_cover_everything_once():
    from footsies_gym_b200.distributed import shard_range
    for total, world in ((1 << 20, 8), (1000, 3), (7, 8), (65536, 2)):
        seen = []
        for r in range(world):
            first, count = shard_range(total, r, world)
            seen += list(range(first, first + count))
        assert seen == list(range(total))


def test_generated_tables_are_current_when_reference_is_present():
    if not os.path.isdir("/root/reference/Assets/Fighter/F00"):
        pytest.skip("reference tree not present (GPU box)")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_frame_data.py"), "--check"],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr


GLOO_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import torch, torch.distributed as dist
from footsies_gym_b200.distributed import all_reduce_stats, shard_range, env_rank_world
rank, local_rank, world = env_rank_world()
dist.init_process_group("gloo", rank=rank, world_size=world)
first, count = shard_range(1001, rank, world)
stats = torch.zeros(16, dtype=torch.int64)
stats[0] = count            # pretend: one episode per env
stats[10] = 10 * count      # env frames
out = all_reduce_stats(stats)
assert out["episodes"] == 1001 and out["env_frames"] == 10010, out
# max-over-ranks timing reduction used by bench.py
t = torch.tensor([float(rank + 1)], dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
assert t.item() == world
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok", first, count)
"""


def test_gloo_world_size_2_stats_all_reduce(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29613", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
