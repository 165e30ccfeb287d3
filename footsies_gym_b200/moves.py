"""Move table mirroring footsies-gym/footsies_gym/moves.py (same names, same member order, same fields),
built from the generated frame data instead of being typed in by hand."""
from dataclasses import dataclass
from enum import Enum

from . import frame_data as _fd


@dataclass
class FootsiesMoveInfo:
    id: int
    duration: int
    startup: int
    active: int
    recovery: int


def _info(action):
    real = [h for h in action["hitboxes"] if not h["proximity"]]
    if real:
        startup = min(h["se"][0] for h in real)
        active = max(h["se"][1] for h in real) - startup + 1
        recovery = action["frameCount"] - startup - active
    else:
        startup = active = recovery = 0
    return FootsiesMoveInfo(action["actionID"], action["frameCount"], startup, active, recovery)


FootsiesMove = Enum("FootsiesMove", {a["actionName"]: _info(a) for a in _fd.ACTIONS})


def _in_recovery(self, frame: int) -> bool:
    return frame >= (self.value.startup + self.value.active)


def _in_active(self, frame: int) -> bool:
    return self.value.startup <= frame < (self.value.startup + self.value.active)


def _in_startup(self, frame: int) -> bool:
    return frame < self.value.startup


FootsiesMove.in_recovery = _in_recovery
FootsiesMove.in_active = _in_active
FootsiesMove.in_startup = _in_startup

# moves.py:41-42
FOOTSIES_MOVE_INDEX_TO_MOVE = list(FootsiesMove)
FOOTSIES_MOVE_ID_TO_INDEX = {move.value.id: i for i, move in enumerate(FOOTSIES_MOVE_INDEX_TO_MOVE)}
