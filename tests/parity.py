"""Field-by-field comparison of the CUDA path (through the C ABI) with the CPU oracle.

Bar: every integer field bit-exact; positions / velocities / observations / rewards compared with ==
(stated tolerance: 0 ulp -- both sides evaluate the same non-contracted fp32 expressions, and the reward
is the fp32 cast of the same float64 value).
"""
import numpy as np

FIGHTER_INT_FIELDS = ["action_id", "action_frame", "hitstun", "guard", "vital", "hit_count", "buffer_id",
                      "reserve_id", "is_input_backward", "is_reserve_prox", "shake", "has_won", "input0",
                      "attack_run"]
FIGHTER_FLOAT_FIELDS = ["pos_x", "velocity_x"]


def _fail(where, name, got, exp):
    bad = np.argwhere(np.asarray(got) != np.asarray(exp))
    i = tuple(bad[0])
    raise AssertionError(
        f"[{where}] {name} mismatch at {i} ({len(bad)} entries differ): kernel={np.asarray(got)[i]!r} "
        f"oracle={np.asarray(exp)[i]!r}")


def _eq(where, name, got, exp):
    got = np.asarray(got)
    exp = np.asarray(exp)
    if got.shape != exp.shape:
        raise AssertionError(f"[{where}] {name} shape {got.shape} vs {exp.shape}")
    if not np.array_equal(got, exp):
        _fail(where, name, got, exp)


def canonical_history(left, right):
    """The kernel keeps, of the Left/Right input history, exactly what dash detection can ever read (Fighter.cs:585-635):
    how long ago the most recent direction-held frame of the last 8 was (`since`), which directions it held and the
    length (capped at 9) of the unbroken run of direction-held frames ending there.  fg_get_state expands that back into
    a bit history; this reduces the oracle's real history (bit i = held i frames ago) to the same canonical form."""
    left = np.asarray(left, dtype=np.uint32) & 0xFFFF
    right = np.asarray(right, dtype=np.uint32) & 0xFFFF
    anyd = left | right
    since = np.full(left.shape, 8, dtype=np.int64)
    for k in range(7, -1, -1):
        since = np.where((anyd >> k) & 1 == 1, k, since)
    cold = since >= 8
    s = np.minimum(since, 7)
    runlen = np.zeros(left.shape, dtype=np.int64)
    alive = ~cold
    for j in range(16):
        pos = s + j
        held = alive & (pos < 16) & (((anyd >> np.minimum(pos, 15)) & 1) == 1) & (runlen < 9)
        runlen = np.where(held, runlen + 1, runlen)
        alive = alive & held
    last_l = (left >> s) & 1
    last_r = (right >> s) & 1
    out_l = np.zeros(left.shape, dtype=np.uint32)
    out_r = np.zeros(left.shape, dtype=np.uint32)
    end = np.where(runlen >= 9, 16, s + runlen)
    for k in range(16):
        inside = (~cold) & (k >= s) & (k < end)
        out_l |= np.where(inside & (last_l == 1), np.uint32(1 << k), np.uint32(0))
        out_r |= np.where(inside & (last_r == 1), np.uint32(1 << k), np.uint32(0))
    return out_l, out_r


def compare_states(kernel_state, oracle_trace, where="", check_rng=True, check_actor=(True, True)):
    """kernel_state: structured array from FootsiesEnv.get_state(); oracle_trace: OracleBatch.trace."""
    ks, ot = kernel_state, oracle_trace
    for f in FIGHTER_INT_FIELDS:
        _eq(where, f"f.{f}", ks["f"][f], ot["f"][f])
    for f in FIGHTER_FLOAT_FIELDS:
        _eq(where, f"f.{f}", ks["f"][f], ot["f"][f])
    exp_l, exp_r = canonical_history(ot["f"]["hist_left"], ot["f"]["hist_right"])
    _eq(where, "f.hist_left (canonical form)", ks["f"]["hist_left"] & 0xFFFF, exp_l)
    _eq(where, "f.hist_right (canonical form)", ks["f"]["hist_right"] & 0xFFFF, exp_r)
    _eq(where, "frame", ks["frame"], ot["frame"])
    _eq(where, "recorded_input", ks["recorded_input"], ot["recorded_input"])
    _eq(where, "done", ks["done"], ot["terminated"])
    for s in (0, 1):
        if check_actor[s]:
            _eq(where, f"actor_input[{s}]", ks["actor_input"][:, s], ot["bot_input"][:, s])
    if check_rng:
        _eq(where, "rng_state", ks["rng_state"], ot["rng_state"])


def compare_outputs(env, oracle_trace, where=""):
    ot = oracle_trace
    _eq(where, "obs", env.obs.cpu().numpy(), ot["obs"])
    _eq(where, "reward", env.reward.cpu().numpy(), ot["reward"])
    _eq(where, "terminated", env.terminated.cpu().numpy().astype(np.int32), ot["terminated"])
    _eq(where, "info_frame", env.info_frame.cpu().numpy(), ot["info_frame"])
    misc = env.info_misc.cpu().numpy().astype(np.int32)
    _eq(where, "info_action", misc[:, 0:2], ot["info_action"])
    _eq(where, "info_hitstun", misc[:, 2:4], ot["info_hitstun"])


def compare_state_and_outputs(env, oracle_trace, where="", check_rng=True):
    compare_states(env.get_state(), oracle_trace, where, check_rng=check_rng)
    compare_outputs(env, oracle_trace, where)


STAT_MAP = {"episodes": "episodes", "p1_wins": "p1_wins", "p2_wins": "p2_wins", "double_ko": "double_ko",
            "episode_frames": "frames", "p1_specials": "p1_specials", "p1_specials_neutral": "p1_specials_neutral",
            "guard_breaks": "guard_breaks", "hits": "hits", "blocks": "blocks"}


def compare_stats(env, oracle_batch, where=""):
    ks, os_ = env.episode_stats(), oracle_batch.stats()
    for k, o in STAT_MAP.items():
        if ks[k] != os_[o]:
            raise AssertionError(f"[{where}] stat {k}: kernel={ks[k]} oracle={os_[o]}")
    if ks["env_frames"] != oracle_batch.frames_simulated():
        raise AssertionError(f"[{where}] env_frames: kernel={ks['env_frames']} oracle={oracle_batch.frames_simulated()}")
