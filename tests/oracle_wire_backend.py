"""Wire-server backend whose engine is the CPU oracle.

TEST INFRASTRUCTURE ONLY (used by tests/ and tests/golden/make_golden.py): lets footsies_gym_b200.wire's protocol
logic be exercised on a machine without a GPU, and lets the reference's own FootsiesEnv record golden vectors over
its real TCP protocol.  The product ships GpuBattleBackend only."""
import json

import oracle_binding as ob


class OracleBattleBackend:
    def __init__(self, seed=0, p2_bot=True):
        self._p2_bot = p2_bot
        self._seed = seed
        # game side only: the oracle's own python half is irrelevant here (autoreset off, delay 0)
        self.orc = ob.OracleBatch(1, p2_bot=p2_bot, autoreset=False, seed=seed)

    @property
    def p2_bot(self):
        return self._p2_bot

    def _state(self):
        t = self.orc.trace[0]
        f1, f2 = t["f"][0], t["f"][1]
        return {
            "p1Vital": int(f1["vital"]), "p2Vital": int(f2["vital"]), "p1Guard": int(f1["guard"]),
            "p2Guard": int(f2["guard"]), "p1Move": int(f1["action_id"]), "p1MoveFrame": int(f1["action_frame"]),
            "p2Move": int(f2["action_id"]), "p2MoveFrame": int(f2["action_frame"]),
            "p1Position": float(f1["pos_x"]), "p2Position": float(f2["pos_x"]), "globalFrame": int(t["frame"]),
            "p1MostRecentAction": int(t["recorded_input"][0]), "p2MostRecentAction": int(t["recorded_input"][1]),
            "p1Hitstun": int(f1["hitstun"]), "p2Hitstun": int(f2["hitstun"]),
        }

    def reset(self):
        self.orc.reset()
        return self._state()

    def step(self, a1, a2):
        self.orc.step([a1], [a2])
        return self._state(), bool(self.orc.trace[0]["battle_over"])

    def seed(self, value):
        self.orc.seed(int(value))

    def save_battle_state(self):
        return json.dumps(self.orc.save_battle_state(0))

    def load_battle_state(self, battle_state_json):
        self.orc.load_battle_state(0, json.loads(battle_state_json))

    def set_p2_bot(self, bot):
        raise NotImplementedError("the oracle backend fixes the P2 actor at construction")

    def close(self):
        pass
