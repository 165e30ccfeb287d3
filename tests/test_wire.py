"""The wire-protocol shim (footsies_gym_b200/wire.py): a minimal agent written against the
reference's protocol description (footsies.py:261-334, 407-456; SocketHelper.cs:48-82; TrainingRemoteControl.cs)
drives the server; every state the "game" sends is compared with a directly driven CPU oracle.

The unmodified reference FootsiesEnv itself is run against this same server by tests/golden/make_golden.py (it
cannot be imported on the GPU box); the goldens it recorded are replayed by test_golden_reference_python.py.
"""
import json
import socket

import numpy as np
import pytest

from footsies_gym_b200.wire import FootsiesWireServer, recv_message, send_message

FIELDS = ("p1Vital", "p2Vital", "p1Guard", "p2Guard", "p1Move", "p1MoveFrame", "p2Move", "p2MoveFrame", "p1Position",
          "p2Position", "globalFrame", "p1MostRecentAction", "p2MostRecentAction", "p1Hitstun", "p2Hitstun")


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


class Agent:
    """What FootsiesEnv does on the wire, without gymnasium."""

    def __init__(self, ports, with_opponent):
        self.p1 = socket.create_connection(("127.0.0.1", ports[0]))
        self.rc = socket.create_connection(("127.0.0.1", ports[1]))
        self.p2 = socket.create_connection(("127.0.0.1", ports[2])) if with_opponent else None
        for s in (self.p1, self.rc, self.p2):
            if s is not None:
                s.settimeout(20)

    def state(self):
        raw = recv_message(self.p1)
        d = json.loads(raw.decode("utf-8"))
        assert tuple(d.keys()) == FIELDS                                   # EnvironmentState.cs field order
        return d

    def act(self, a1, a2=None):
        self.p1.sendall(bytes([a1 & 1, (a1 >> 1) & 1, (a1 >> 2) & 1]))      # footsies.py:323-334
        if self.p2 is not None and a2 is not None:
            self.p2.sendall(bytes([a2 & 1, (a2 >> 1) & 1, (a2 >> 2) & 1]))

    def command(self, cmd, value=""):
        send_message(self.rc, json.dumps({"command": cmd, "value": value}).encode("utf-8"))

    def save(self):
        self.command(2)
        return recv_message(self.rc).decode("utf-8")

    def close(self):
        for s in (self.p1, self.rc, self.p2):
            if s is not None:
                s.close()


def expected_state(orc):
    t = orc.trace[0]
    f1, f2 = t["f"][0], t["f"][1]
    return {"p1Vital": int(f1["vital"]), "p2Vital": int(f2["vital"]), "p1Guard": int(f1["guard"]),
            "p2Guard": int(f2["guard"]), "p1Move": int(f1["action_id"]), "p1MoveFrame": int(f1["action_frame"]),
            "p2Move": int(f2["action_id"]), "p2MoveFrame": int(f2["action_frame"]), "p1Position": float(f1["pos_x"]),
            "p2Position": float(f2["pos_x"]), "globalFrame": int(t["frame"]),
            "p1MostRecentAction": int(t["recorded_input"][0]), "p2MostRecentAction": int(t["recorded_input"][1]),
            "p1Hitstun": int(f1["hitstun"]), "p2Hitstun": int(f2["hitstun"])}


def drive(oracle, backend, p2_remote, steps, rng_seed):
    """Runs the scripted session; returns the number of compared states."""
    ports = free_ports(3)
    server = FootsiesWireServer(backend, ports[0], ports[1], ports[2] if p2_remote else None)
    server.start()
    agent = Agent(ports, p2_remote)
    orc = oracle.OracleBatch(1, p2_bot=not p2_remote, autoreset=False, seed=0)
    rng = np.random.default_rng(rng_seed)
    compared = 0
    try:
        orc.reset()
        assert agent.state() == expected_state(orc)                        # the game starts a round by itself
        agent.command(5, "1234")                                           # SEED
        orc.seed(1234)
        agent.command(1)                                                   # RESET
        orc.reset()
        assert agent.state() == expected_state(orc)
        saved, saved_at = None, None
        sticky = 0
        for t in range(steps):
            if rng.random() < 0.2:
                sticky = int(rng.integers(0, 8))
            a1 = sticky if rng.random() < 0.7 else int(rng.integers(0, 8))
            a2 = int(rng.integers(0, 8)) if p2_remote else 0
            agent.act(a1, a2 if p2_remote else None)
            orc.step([a1], [a2])
            got = agent.state()
            assert got == expected_state(orc), (t, got, expected_state(orc))
            compared += 1
            if orc.trace[0]["battle_over"]:
                orc.reset()
                assert agent.state() == expected_state(orc)                # automatic restart
                compared += 1
            elif p2_remote and t in (60, 200):                             # STATE_SAVE then, later, STATE_LOAD
                saved = agent.save()
                d = json.loads(saved)
                assert d["frameCount"] == expected_state(orc)["globalFrame"]
                assert d["p1State"]["currentActionID"] == expected_state(orc)["p1Move"]
                assert len(d["p1State"]["input"]) == 180 and d["p2State"]["isFaceRight"] is False
                saved_at = orc.save_battle_state(0)
            elif p2_remote and t in (90, 230) and saved is not None:
                agent.command(3, saved)
                # the oracle loads what the backend would have kept of it: for the oracle backend the state as is,
                # for the GPU backend the compact form -- both continue identically (test_battle_state.py)
                orc.load_battle_state(0, saved_at)
                saved = None
    finally:
        agent.close()
        server.stop()
        server.join(timeout=5)
    return compared


@pytest.mark.parametrize("p2_remote", [False, True])
def test_wire_protocol_with_oracle_backend(oracle, p2_remote):
    from oracle_wire_backend import OracleBattleBackend
    n = drive(oracle, OracleBattleBackend(seed=0, p2_bot=not p2_remote), p2_remote, steps=600, rng_seed=3)
    assert n >= 600


@pytest.mark.gpu
@pytest.mark.parametrize("p2_remote", [False, True])
def test_wire_protocol_with_gpu_backend(oracle, p2_remote):
    from footsies_gym_b200.wire import GpuBattleBackend
    backend = GpuBattleBackend(device="cuda:0", seed=0, p2_bot=not p2_remote)
    try:
        n = drive(oracle, backend, p2_remote, steps=1500, rng_seed=4)
    finally:
        backend.close()
    assert n >= 1500
