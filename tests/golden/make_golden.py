#!/usr/bin/env python3
"""Generate golden vectors with the REFERENCE's own, unmodified Python code.

Runs only in the authoring container (needs /root/reference).  It imports the reference's
`footsies_gym.envs.footsies.FootsiesEnv` (gymnasium is not installed, so a minimal stand-in module is put in
sys.modules first -- the reference code itself is untouched) and lets it talk over its real TCP protocol
(footsies.py:261-334, 407-456; SocketHelper.cs:48-82; TrainingRemoteControl.cs:18-107) to a small game server
whose battle engine is the CPU oracle.  Everything the reference computes in Python -- observation dict,
move-index mapping, DEAD->STAND remap, move_frame simplification, info, dense / sparse reward, termination,
frame_delay queue, reset semantics -- is recorded as the expected answer.

What this pins: the Python half of the path and the EnvironmentState field mapping
(a15).  What it cannot pin: the C# engine itself (the oracle plays the game's part here).

Output: tests/golden/ref_python_<name>.npz, one per scenario:
  ops        int32 [M, 4]  what the game did, in order: (kind, a, b, expect_index)
                           kind 0 = fight frame with P1 action a, P2 action b; 1 = round start (reset);
                           2 = Random.InitState(a)
  exp_*      the reference's outputs for every reset() / step() call, indexed by expect_index
"""
import json
import os
import socket
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF_PY = "/root/reference/footsies-gym"


def install_gymnasium_stub():
    from footsies_gym_b200 import spaces as sp
    gym = types.ModuleType("gymnasium")

    class Env:
        def reset(self, *, seed=None, options=None):
            return None

    class Wrapper:                      # gymnasium.Wrapper semantics: forward everything to the wrapped env
        def __init__(self, env):
            self.env = env
            self.observation_space = getattr(env, "observation_space", None)
            self.action_space = getattr(env, "action_space", None)

        def __getattr__(self, name):
            if name.startswith("_"):
                raise AttributeError(name)
            return getattr(self.env, name)

        def reset(self, *, seed=None, options=None):
            return self.env.reset(seed=seed, options=options)

        def step(self, action):
            return self.env.step(action)

    class ObservationWrapper(Wrapper):  # gymnasium.ObservationWrapper: observation() applied to reset and step
        def reset(self, *, seed=None, options=None):
            obs, info = self.env.reset(seed=seed, options=options)
            return self.observation(obs), info

        def step(self, action):
            obs, reward, terminated, truncated, info = self.env.step(action)
            return self.observation(obs), reward, terminated, truncated, info

    class ActionWrapper(Wrapper):       # gymnasium.ActionWrapper: action() applied before step
        def step(self, action):
            return self.env.step(self.action(action))

    gym.Env = Env
    gym.Wrapper, gym.ObservationWrapper, gym.ActionWrapper = Wrapper, ObservationWrapper, ActionWrapper
    spaces = types.ModuleType("gymnasium.spaces")
    spaces.Dict, spaces.MultiDiscrete, spaces.Box = sp.Dict, sp.MultiDiscrete, sp.Box
    spaces.MultiBinary, spaces.Discrete = sp.MultiBinary, sp.Discrete
    utils = types.ModuleType("gymnasium.spaces.utils")
    utils.unflatten = lambda space, x: x
    spaces.utils = utils
    spaces.Space = object
    envs = types.ModuleType("gymnasium.envs")
    reg = types.ModuleType("gymnasium.envs.registration")
    reg.register = lambda **kw: None
    envs.registration = reg
    gym.spaces, gym.envs = spaces, envs
    for name, mod in (("gymnasium", gym), ("gymnasium.spaces", spaces), ("gymnasium.spaces.utils", utils),
                      ("gymnasium.envs", envs), ("gymnasium.envs.registration", reg)):
        sys.modules[name] = mod


def free_ports(k):
    socks, ports = [], []
    for _ in range(k):
        s = socket.socket()
        s.bind(("127.0.0.1", 0))
        ports.append(s.getsockname()[1])
        socks.append(s)
    for s in socks:
        s.close()
    return ports


class OracleGameServer:
    """The product's wire server (footsies_gym_b200.wire.FootsiesWireServer) playing the Unity game's role, with
    the CPU oracle as its battle engine; records what the game did as (kind, a, b) ops."""

    def __init__(self, ports, p2_remote, seed0=0):
        from footsies_gym_b200.wire import FootsiesWireServer
        from oracle_wire_backend import OracleBattleBackend
        self.backend = OracleBattleBackend(seed=seed0, p2_bot=not p2_remote)
        self.ops = []
        self.server = FootsiesWireServer(self.backend, ports[0], ports[1], ports[2] if p2_remote else None,
                                         on_op=self._on_op)

    def _on_op(self, kind, *args):
        if kind == "frame":
            self.ops.append((0, args[0], args[1]))
        elif kind == "round_start":
            self.ops.append((1, 0, 0))
        elif kind == "seed":
            self.ops.append((2, args[0], 0))

    def start(self):
        self.server.start()

    def stop(self):
        self.server.stop()

    def join(self, timeout=None):
        self.server.join(timeout)


def flat_obs(obs):
    return [obs["guard"][0], obs["guard"][1], obs["move"][0], obs["move"][1],
            obs["move_frame"][0], obs["move_frame"][1], obs["position"][0], obs["position"][1]]


def mask3(t):
    return int(t[0]) | int(t[1]) << 1 | int(t[2]) << 2


def run_scenario(name, *, dense, frame_delay, p2_remote, n_calls, rng_seed, seeds_at_reset=True):
    from footsies_gym.envs.footsies import FootsiesEnv   # the reference class, unmodified
    ports = free_ports(3)
    server = OracleGameServer(ports, p2_remote)
    server.start()
    rng = np.random.default_rng(rng_seed)
    p2_tape = []

    def opponent(obs, info):
        a = int(rng.integers(0, 8))
        p2_tape.append(a)
        return (a & 1 != 0, a & 2 != 0, a & 4 != 0)

    env = FootsiesEnv(frame_delay=frame_delay, game_address="127.0.0.1", game_port=ports[0],
                      remote_control_port=ports[1], opponent_port=ports[2], skip_instancing=True,
                      sync_mode="synced_non_blocking", dense_reward=dense,
                      opponent=opponent if p2_remote else None)
    exp = {k: [] for k in ("kind", "obs", "reward", "terminated", "truncated", "frame", "action", "hitstun")}

    def record(kind, obs, reward, terminated, truncated, info):
        exp["kind"].append(kind)
        exp["obs"].append(flat_obs(obs))
        exp["reward"].append(float(reward))
        exp["terminated"].append(int(terminated))
        exp["truncated"].append(int(truncated))
        exp["frame"].append(int(info["frame"]))
        exp["action"].append([mask3(info["p1_action"]), mask3(info["p2_action"])])
        exp["hitstun"].append([int(info["p1_hitstun"]), int(info["p2_hitstun"])])
        for k in ("guard", "move", "move_frame", "position"):   # info carries a copy of obs (footsies.py:378-379)
            assert tuple(info[k]) == tuple(obs[k])

    # client-side log of which expectation belongs to which game op
    call_ops = []
    episode = 0
    obs, info = env.reset(seed=1000)
    record(1, obs, 0.0, False, False, info)
    call_ops.append("reset")
    steps_in_episode = 0
    sticky = 0
    for _ in range(n_calls):
        if rng.random() < 0.15:
            sticky = int(rng.integers(0, 8))
        a = sticky if rng.random() < 0.7 else int(rng.integers(0, 8))
        obs, reward, terminated, truncated, info = env.step((a & 1 != 0, a & 2 != 0, a & 4 != 0))
        record(0, obs, reward, terminated, truncated, info)
        call_ops.append("step")
        steps_in_episode += 1
        force = (not terminated) and steps_in_episode > 40 and rng.random() < 0.004
        if terminated or force:
            episode += 1
            seed = 2000 + episode if (seeds_at_reset and episode % 3 == 0) else None
            obs, info = env.reset(seed=seed)
            record(1, obs, 0.0, False, False, info)
            call_ops.append("reset")
            steps_in_episode = 0
    server.stop()
    env.close()
    server.join(timeout=5)

    # attach expectation indices to the game ops: every step op <-> one step() call in order; every reset()
    # call <-> the LAST round start before the next fight frame (earlier ones were never observed)
    ops = server.ops
    step_calls = [i for i, k in enumerate(call_ops) if k == "step"]
    reset_calls = [i for i, k in enumerate(call_ops) if k == "reset"]
    out = np.full((len(ops), 4), -1, dtype=np.int32)
    si = 0
    for j, (kind, a, b) in enumerate(ops):
        out[j, :3] = (kind, a, b)
        if kind == 0:
            out[j, 3] = step_calls[si]
            si += 1
    assert si == len(step_calls)
    # walk the client calls and the ops together to place the reset expectations
    j = 0
    last_round_start = None
    for i, k in enumerate(call_ops):
        if k == "step":
            while ops[j][0] != 0:
                if ops[j][0] == 1:
                    last_round_start = j
                j += 1
            j += 1
        else:
            # the observed round start is the last kind-1 op before the next kind-0 op (or the end)
            jj = j
            cand = None
            while jj < len(ops) and ops[jj][0] != 0:
                if ops[jj][0] == 1:
                    cand = jj
                jj += 1
            if cand is None:
                cand = last_round_start
            out[cand, 3] = i
    path = os.path.join(HERE, f"ref_python_{name}.npz")
    np.savez_compressed(
        path, ops=out, exp_kind=np.array(exp["kind"], np.int8), exp_obs=np.array(exp["obs"], np.float64),
        exp_reward=np.array(exp["reward"], np.float64), exp_terminated=np.array(exp["terminated"], np.int8),
        exp_truncated=np.array(exp["truncated"], np.int8), exp_frame=np.array(exp["frame"], np.int32),
        exp_action=np.array(exp["action"], np.int8), exp_hitstun=np.array(exp["hitstun"], np.int8),
        p2_tape=np.array(p2_tape, np.int8),
        config=np.array([int(dense), frame_delay, int(p2_remote), 0], np.int32))
    n_ep = int(np.sum(np.array(exp["terminated"])))
    print(f"{name}: {len(ops)} game ops, {len(call_ops)} API calls, {n_ep} terminations -> {os.path.relpath(path, ROOT)}")


def run_wrapper_scenario(name, chain, n_calls, rng_seed):
    """Reference wrappers (footsies_gym/wrappers/*.py, unmodified) stacked on the reference FootsiesEnv; the
    recorded API-level calls are replayed through the batched wrappers of footsies_gym_b200.wrappers."""
    from footsies_gym.envs.footsies import FootsiesEnv
    from footsies_gym.wrappers import (FootsiesActionCombinationsDiscretized, FootsiesFrameSkipped,
                                       FootsiesNormalized, FootsiesStatistics)
    ports = free_ports(3)
    server = OracleGameServer(ports, p2_remote=False)
    server.start()
    env = FootsiesEnv(game_address="127.0.0.1", game_port=ports[0], remote_control_port=ports[1],
                      opponent_port=ports[2], skip_instancing=True, sync_mode="synced_non_blocking")
    stats = None
    for w in chain:
        if w == "normalized":
            env = FootsiesNormalized(env)
        elif w == "frame_skipped":
            env = FootsiesFrameSkipped(env)
        elif w == "discretized":
            env = FootsiesActionCombinationsDiscretized(env)
        elif w == "statistics":
            env = stats = FootsiesStatistics(env)
    rng = np.random.default_rng(rng_seed)
    calls, obs_l, rew_l, term_l, frame_l = [], [], [], [], []

    def rec(kind, a, obs, reward, terminated, info):
        calls.append((kind, a))
        mf = obs["move_frame"]
        mf = list(mf) if isinstance(mf, (tuple, list)) else [float(mf), float("nan")]
        obs_l.append([obs["guard"][0], obs["guard"][1], obs["move"][0], obs["move"][1], mf[0], mf[1],
                      obs["position"][0], obs["position"][1]])
        rew_l.append(float(reward))
        term_l.append(int(terminated))
        frame_l.append(int(info["frame"]))

    obs, info = env.reset(seed=None, options=None)
    rec(1, 0, obs, 0.0, False, info)
    sticky = 0
    for _ in range(n_calls):
        if rng.random() < 0.2:
            sticky = int(rng.integers(0, 8))
        a = sticky if rng.random() < 0.7 else int(rng.integers(0, 8))
        act = a if "discretized" in chain else (a & 1 != 0, a & 2 != 0, a & 4 != 0)
        obs, reward, terminated, truncated, info = env.step(act)
        rec(0, a, obs, reward, terminated, info)
        if terminated:
            obs, info = env.reset(seed=None, options=None)
            rec(1, 0, obs, 0.0, False, info)
    server.stop()
    env.close()
    server.join(timeout=5)
    extra = {}
    if stats is not None:
        extra["special_moves_per_episode"] = np.array(stats.metric_special_moves_per_episode, np.int32)
        extra["special_moves_from_neutral_per_episode"] = np.array(stats.metric_special_moves_from_neutral_per_episode, np.int32)
    path = os.path.join(HERE, f"ref_wrappers_{name}.npz")
    np.savez_compressed(path, calls=np.array(calls, np.int32), exp_obs=np.array(obs_l, np.float64),
                        exp_reward=np.array(rew_l, np.float64), exp_terminated=np.array(term_l, np.int8),
                        exp_frame=np.array(frame_l, np.int32), chain=np.array(chain), **extra)
    print(f"{name}: {len(calls)} API calls, {len(server.ops)} game frames/resets, {int(np.sum(term_l))} terminations "
          f"-> {os.path.relpath(path, ROOT)}")


def run_battle_state_fixture():
    """STATE_SAVE / STATE_LOAD through the reference's own client code and dataclasses (footsies.py:407-444,
    state.py:78-137): the reference FootsiesEnv asks the wire server for battle states, its FootsiesBattleState
    parses and re-serialises them, and its FootsiesState.from_battle_state gives the summary view.  The fixture pins
    the JSON schema and field order our state.py must reproduce."""
    import dataclasses
    from footsies_gym.envs.footsies import FootsiesEnv
    from footsies_gym.state import FootsiesBattleState as RefBattleState, FootsiesFighterState as RefFighterState
    from footsies_gym.state import FootsiesState as RefState
    ports = free_ports(3)
    server = OracleGameServer(ports, p2_remote=False)
    server.start()
    env = FootsiesEnv(game_address="127.0.0.1", game_port=ports[0], remote_control_port=ports[1],
                      opponent_port=ports[2], skip_instancing=True, sync_mode="synced_non_blocking")
    rng = np.random.default_rng(11)
    env.reset(seed=5)
    cases = []
    sticky = 0
    for t in range(400):
        if rng.random() < 0.2:
            sticky = int(rng.integers(0, 8))
        obs, reward, terminated, truncated, info = env.step((sticky & 1 != 0, sticky & 2 != 0, sticky & 4 != 0))
        if terminated:
            env.reset()
        elif t % 40 == 17:
            bs = env.save_battle_state()                       # reference client + reference dataclass parsing
            st = RefState.from_battle_state(bs)
            d = dataclasses.asdict(st)
            d["p1MostRecentAction"] = list(d["p1MostRecentAction"])
            d["p2MostRecentAction"] = list(d["p2MostRecentAction"])
            cases.append({"reference_json": bs.json(), "reference_footsies_state": d})
            # a load of the state just saved must leave the game where it is (round trip through the reference)
            env.load_battle_state(bs)
            again = env.save_battle_state()
            assert again.json() == bs.json()
    server.stop()
    env.close()
    server.join(timeout=5)
    path = os.path.join(HERE, "ref_battle_state.json")
    with open(path, "w") as f:
        json.dump({"fighter_fields": [x.name for x in dataclasses.fields(RefFighterState)],
                   "battle_fields": [x.name for x in dataclasses.fields(RefBattleState)], "cases": cases}, f)
    print(f"battle_state: {len(cases)} saved states -> {os.path.relpath(path, ROOT)}")


def main():
    install_gymnasium_stub()
    sys.path.insert(0, REF_PY)
    run_battle_state_fixture()
    run_scenario("dense_bot", dense=True, frame_delay=0, p2_remote=False, n_calls=6000, rng_seed=1)
    run_scenario("sparse_bot", dense=False, frame_delay=0, p2_remote=False, n_calls=3000, rng_seed=2)
    run_scenario("dense_delay3_bot", dense=True, frame_delay=3, p2_remote=False, n_calls=3000, rng_seed=3)
    run_scenario("dense_remote_p2", dense=True, frame_delay=0, p2_remote=True, n_calls=4000, rng_seed=4)
    run_wrapper_scenario("normalized_frameskipped", ["normalized", "frame_skipped"], n_calls=2500, rng_seed=5)
    run_wrapper_scenario("discretized_statistics", ["discretized", "statistics"], n_calls=4000, rng_seed=6)
    run_wrapper_scenario("normalized", ["normalized"], n_calls=1500, rng_seed=7)


if __name__ == "__main__":
    main()
