"""Invariants the REFERENCE ITSELF states about the path (it ships no tests; these are its specification):

  * every observation lies inside FootsiesEnv.observation_space (footsies.py:153-168): guard in 0..3, move index in
    0..14 (WIN / DEAD never observed), move_frame in [0, 55], position in [-4.6, 4.6];
  * reward_range (-1, 1) and the dense-reward docstring (footsies.py:91-92, 388-405): the rewards of an episode sum to
    exactly +-1 (to 1e-6: Python float64 arithmetic, delivered as float32), sparse rewards are 0 until the +-1 at the end;
  * README.md:36-46 "You can block opponent attack up to three times. After that, every attack will cause guard break":
    the guard bar only ever drops, by one per connect, is full at a round start, a block with bar left costs one and never
    breaks, and the guard-break stun (30 frames, F00_AttackDataContainer.asset) only ever appears on an already-empty bar;
  * truncation never happens and the first observation of an episode is frame -1 (footsies.py:497-500, 570).

Checked with torch ops on whatever device the tensors live on, so the B200 tests run them over a million battles every
step; the CPU suite runs them over the oracle and over the transliterated reference engine."""
import numpy as np
import torch

GUARD_MOVES = (10, 11, 12)      # GUARD_M, GUARD_STAND, GUARD_CROUCH in moves.py order
POSITION_BOUND = float(np.float32(4.6))


class ReferenceInvariants:
    def __init__(self, n, device, dense=True):
        self.n, self.dense = n, dense
        self.ret = torch.zeros(n, dtype=torch.float64, device=device)
        self.prev_guard = None
        self.prev_stun = None
        self.fresh = torch.ones(n, dtype=torch.bool, device=device)     # next observation is a reset observation
        self.episodes = 0
        self.breaks = 0
        self.blocks = 0

    def reset(self, obs, info_frame, info_misc):
        assert bool((info_frame == -1).all()) and bool((obs[:, 0:2] == 3).all())
        self.prev_guard = obs[:, 0:2].clone()
        self.prev_stun = info_misc[:, 2:4].to(torch.int32)
        self.ret.zero_()
        self.fresh.zero_()

    def update(self, obs, reward, terminated, info_frame, info_misc, where=""):
        """Call after every step.  `terminated` envs are expected to be reset by their next step (autoreset) or frozen."""
        guard, move, mframe, pos = obs[:, 0:2], obs[:, 2:4], obs[:, 4:6], obs[:, 6:8]
        stun = info_misc[:, 2:4].to(torch.int32)
        was_reset = info_frame == -1
        # ---- observation_space (footsies.py:153-168)
        assert bool(((guard >= 0) & (guard <= 3) & (guard == guard.round())).all()), f"{where}: guard outside MultiDiscrete([4, 4])"
        assert bool(((move >= 0) & (move <= 14) & (move == move.round())).all()), f"{where}: move outside MultiDiscrete([15, 15])"
        assert bool(((mframe >= 0) & (mframe <= 55)).all()), f"{where}: move_frame outside Box(0, 55)"
        assert float(pos.abs().max()) <= POSITION_BOUND, f"{where}: |position| = {float(pos.abs().max())!r} outside Box(-4.6, 4.6)"
        # ---- reward (footsies.py:91-92, 382-405) and reward_range
        r = reward.double()
        # (a dense step reward may leave reward_range in the reference itself: winning after -0.6 pays +1.6; the SUM is +-1)
        assert float(r.abs().max()) <= (1.9 if self.dense else 1.0) + 1e-6, f"{where}: step reward out of bounds"
        assert bool((r[was_reset] == 0).all()), f"{where}: a reset observation carries a reward"
        self.ret += r
        if not self.dense:
            assert bool((r[~terminated] == 0).all()), f"{where}: sparse reward before the end"
        done_ret = self.ret[terminated]
        assert bool(((done_ret.abs() - 1.0).abs() <= 1e-6).all()), \
            f"{where}: an episode's rewards do not sum to +-1: {done_ret[((done_ret.abs() - 1.0).abs() > 1e-6)][:4].tolist()}"
        self.episodes += int(terminated.sum())
        self.ret[terminated] = 0.0
        # ---- guard bar (README.md:36-46)
        assert bool((guard[was_reset] == 3).all()), f"{where}: round does not start with a full guard bar"
        running = ~was_reset
        drop = (self.prev_guard - guard)[running]
        assert bool(((drop == 0) | (drop == 1)).all()), f"{where}: guard bar rose or dropped by more than one"
        fresh_stun = (stun > self.prev_stun) & running.unsqueeze(1)
        in_guard = (move == 10) | (move == 11) | (move == 12)
        # guard stun is 12 / 15 frames, guard-break stun 30 (a frame-skipping env may show it a few ticks later);
        # the attacker is put into the same stun (BattleCore.cs:576-578) but is not in a guard action
        broke = fresh_stun & (stun > 15) & in_guard
        assert bool((guard[broke] == 0).all() and (self.prev_guard[broke] == 0).all()), \
            f"{where}: guard break with guard bar left (needs three earlier connects)"
        blocked = fresh_stun & in_guard & (stun <= 15)
        assert bool((self.prev_guard[blocked] >= 1).all() and ((self.prev_guard - guard)[blocked] == 1).all()), \
            f"{where}: a block on an empty bar did not break / a block did not cost one guard point"
        entered_break = (move == 13)
        assert bool((guard[entered_break] == 0).all()), f"{where}: GUARD_BREAK action with guard bar left"
        self.breaks += int(broke.sum())
        self.blocks += int(blocked.sum())
        self.prev_guard = guard.clone()
        self.prev_stun = stun
