"""The game's side of the reference's TCP protocol, served by the GPU simulator.

Lets the UNMODIFIED reference `footsies_gym.envs.footsies.FootsiesEnv(skip_instancing=True, game_port=...,
remote_control_port=..., opponent_port=...)` talk to this simulator as if it were the Unity build:

  * agent socket(s) (TrainingRemoteActor.cs:45-116, footsies.py:308-334): the agent sends 3 raw bytes
    [left, right, attack] per frame; the game answers with a 4-byte big-endian length + the UTF-8
    `JsonUtility.ToJson(EnvironmentState)` (SocketHelper.cs:48-82, EnvironmentState.cs:12-26);
  * remote-control socket (TrainingRemoteControl.cs:18-107, footsies.py:407-456): length-prefixed
    {"command": 0..5, "value": str}: RESET, STATE_SAVE (answered with the BattleState JSON), STATE_LOAD, P2_BOT, SEED.

This is a compatibility shim for single battles, not a fast path: the batched FootsiesEnv is the product API.
The engine behind the sockets is a duck-typed backend; GpuBattleBackend is the one shipped (one battle on the
GPU through the C ABI -- there is no CPU engine in this package).
"""
import json
import select
import socket
import struct
import threading
from typing import Optional

CMD_NONE, CMD_RESET, CMD_STATE_SAVE, CMD_STATE_LOAD, CMD_P2_BOT, CMD_SEED = range(6)


def environment_state_json(s: dict) -> bytes:
    """EnvironmentState in the field order of EnvironmentState.cs:12-26 (what JsonUtility.ToJson emits)."""
    ordered = {k: s[k] for k in (
        "p1Vital", "p2Vital", "p1Guard", "p2Guard", "p1Move", "p1MoveFrame", "p2Move", "p2MoveFrame", "p1Position",
        "p2Position", "globalFrame", "p1MostRecentAction", "p2MostRecentAction", "p1Hitstun", "p2Hitstun")}
    return json.dumps(ordered, separators=(",", ":")).encode("utf-8")


def send_message(sock: socket.socket, payload: bytes):
    sock.sendall(struct.pack("!I", len(payload)) + payload)          # SocketHelper.cs:70-82


def recv_exact(sock: socket.socket, n: int) -> bytes:
    buf = b""
    while len(buf) < n:
        chunk = sock.recv(n - len(buf))
        if not chunk:
            raise ConnectionError("peer closed the connection")
        buf += chunk
    return buf


def recv_message(sock: socket.socket) -> bytes:
    return recv_exact(sock, struct.unpack("!I", recv_exact(sock, 4))[0])


def action_mask(three_bytes: bytes) -> int:
    """3 raw bytes (left, right, attack) -> InputDefine bitmask (TrainingRemoteActor.cs:113-116)."""
    return (1 if three_bytes[0] else 0) | (2 if three_bytes[1] else 0) | (4 if three_bytes[2] else 0)


class GpuBattleBackend:
    """One battle on the GPU behind the wire server.  P2 is the in-game bot or a remote actor (toggled by P2_BOT)."""

    def __init__(self, device="cuda:0", seed: Optional[int] = 0, p2_bot: bool = True):
        from .env import FootsiesEnv
        self._device, self._seed, self._p2_bot = device, seed, p2_bot
        self._env = FootsiesEnv(num_envs=1, device=device, opponent=None if p2_bot else "remote", autoreset=False,
                                seed=seed)

    @property
    def p2_bot(self):
        return self._p2_bot

    def _state(self):
        r = self._env.get_state(0, 1)[0]
        f1, f2 = r["f"][0], r["f"][1]
        return {
            "p1Vital": int(f1["vital"]), "p2Vital": int(f2["vital"]), "p1Guard": int(f1["guard"]),
            "p2Guard": int(f2["guard"]), "p1Move": int(f1["action_id"]), "p1MoveFrame": int(f1["action_frame"]),
            "p2Move": int(f2["action_id"]), "p2MoveFrame": int(f2["action_frame"]),
            "p1Position": float(f1["pos_x"]), "p2Position": float(f2["pos_x"]), "globalFrame": int(r["frame"]),
            "p1MostRecentAction": int(r["recorded_input"][0]), "p2MostRecentAction": int(r["recorded_input"][1]),
            "p1Hitstun": int(f1["hitstun"]), "p2Hitstun": int(f2["hitstun"]),
        }

    def reset(self) -> dict:
        self._env.hard_reset()
        return self._state()

    def step(self, a1: int, a2: int):
        import torch
        p1 = torch.tensor([a1], dtype=torch.uint8)
        if self._p2_bot:
            _, _, terminated, _, _ = self._env.step(p1)
        else:
            _, _, terminated, _, _ = self._env.step(p1, torch.tensor([a2], dtype=torch.uint8))
        return self._state(), bool(terminated.item())

    def seed(self, value: int):
        self._env.seed(int(value))

    def save_battle_state(self) -> str:
        return self._env.save_battle_state(0).json()

    def load_battle_state(self, battle_state_json: str):
        self._env.load_battle_state(battle_state_json, 0)

    def set_p2_bot(self, bot: bool):
        if bot != self._p2_bot:
            self._p2_bot = bot
            self._env.set_opponent(None if bot else (lambda obs, info: None))

    def close(self):
        self._env.close()


class FootsiesWireServer(threading.Thread):
    """Plays the Unity game's part on the wire for ONE agent (+ optionally one remote opponent).

    backend: object with reset() -> state dict, step(a1, a2) -> (state dict, battle_over), seed(int),
             save_battle_state() -> json str, load_battle_state(json str), set_p2_bot(bool), p2_bot (bool)
    Ports are bound on construction; start() then accepts P1, the remote control and (if opponent_port is given)
    the opponent, in the order FootsiesEnv._connect_to_game connects (footsies.py:276-290).
    """

    def __init__(self, backend, game_port: int, remote_control_port: int, opponent_port: Optional[int] = None,
                 address: str = "127.0.0.1", p2_no_state: bool = True, on_op=None):
        super().__init__(daemon=True)
        self.backend = backend
        self.p2_no_state = p2_no_state
        self.on_op = on_op or (lambda *op: None)        # observer hook: ("frame", a1, a2) / ("round_start",) / ("seed", v) ...
        self._stop_flag = threading.Event()
        self._listeners = []
        for port in (game_port, remote_control_port) + ((opponent_port,) if opponent_port is not None else ()):
            ls = socket.socket()
            ls.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
            ls.bind((address, port))
            ls.listen(1)
            self._listeners.append(ls)
        self.error = None

    def stop(self):
        self._stop_flag.set()

    # ---- game events ----
    def _publish(self, state):
        payload = environment_state_json(state)
        send_message(self._p1, payload)                               # TrainingRemoteActor.UpdateCurrentState
        if self._p2 is not None and not self.p2_no_state and not self.backend.p2_bot:
            send_message(self._p2, payload)

    def _round_start(self):
        state = self.backend.reset()                                  # Stop -> Intro -> Fight (BattleCore.cs:176-200)
        self.on_op("round_start")
        self._publish(state)

    def _command(self):
        msg = json.loads(recv_message(self._rc).decode("utf-8"))     # TrainingRemoteControl.ProcessCommand
        cmd, value = int(msg["command"]), msg.get("value", "")
        if cmd == CMD_RESET:
            self._round_start()
        elif cmd == CMD_STATE_SAVE:
            send_message(self._rc, self.backend.save_battle_state().encode("utf-8"))
            self.on_op("state_save")
        elif cmd == CMD_STATE_LOAD:
            self.backend.load_battle_state(value)
            self.on_op("state_load")
        elif cmd == CMD_P2_BOT:
            self.backend.set_p2_bot(str(value).lower() == "true")
            self.on_op("p2_bot", self.backend.p2_bot)
        elif cmd == CMD_SEED:
            self.backend.seed(int(value))
            self.on_op("seed", int(value))

    def run(self):
        try:
            self._p1, _ = self._listeners[0].accept()
            self._rc, _ = self._listeners[1].accept()
            self._p2 = None
            if len(self._listeners) > 2:
                self._p2, _ = self._listeners[2].accept()
            self._round_start()
            while not self._stop_flag.is_set():
                ready, _, _ = select.select([self._p1, self._rc], [], [], 0.1)
                if self._rc in ready:                                 # commands are handled first (BattleCore.cs:140-174)
                    self._command()
                    continue
                if self._p1 in ready:
                    a1 = action_mask(recv_exact(self._p1, 3))
                    a2 = 0
                    if self._p2 is not None and not self.backend.p2_bot:
                        a2 = action_mask(recv_exact(self._p2, 3))
                    state, battle_over = self.backend.step(a1, a2)
                    self.on_op("frame", a1, a2)
                    self._publish(state)
                    if battle_over:                                   # the game restarts by itself (BattleCore.cs:212-217)
                        self._round_start()
        except (ConnectionError, OSError) as e:                       # the agent closed its sockets: quit like the game
            self.error = e
        finally:
            for s in [getattr(self, "_p1", None), getattr(self, "_rc", None), getattr(self, "_p2", None)] + self._listeners:
                try:
                    if s is not None:
                        s.close()
                except OSError:
                    pass


def serve(game_port=11000, remote_control_port=11002, opponent_port=None, device="cuda:0", seed=0):
    """Blocking helper: `python -m footsies_gym_b200.wire` then point the reference FootsiesEnv at the ports."""
    backend = GpuBattleBackend(device=device, seed=seed, p2_bot=opponent_port is None)
    server = FootsiesWireServer(backend, game_port, remote_control_port, opponent_port)
    server.start()
    try:
        server.join()
    finally:
        backend.close()


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser(description="Serve the FOOTSIES game protocol from the GPU simulator")
    ap.add_argument("--game-port", type=int, default=11000)            # footsies.py:40
    ap.add_argument("--remote-control-port", type=int, default=11002)  # footsies.py:46
    ap.add_argument("--opponent-port", type=int, default=None)         # footsies.py:45 (11001) when P2 is remote
    ap.add_argument("--device", default="cuda:0")
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    serve(a.game_port, a.remote_control_port, a.opponent_port, a.device, a.seed)
