"""footsies_gym_b200 -- B200-native batched FOOTSIES simulator behind the FootsiesEnv API.

Importing the package never touches CUDA; constructing a FootsiesEnv does, and fails loudly without a
GPU or without the in-tree library (python -m footsies_gym_b200.build).
"""
from .env import FootsiesEnv, FootsiesGameClosedError
from .state import FootsiesBattleState, FootsiesFighterState, FootsiesState
from .moves import FOOTSIES_MOVE_ID_TO_INDEX, FOOTSIES_MOVE_INDEX_TO_MOVE, FootsiesMove, FootsiesMoveInfo

__all__ = ["FootsiesEnv", "FootsiesGameClosedError", "FootsiesState", "FootsiesBattleState", "FootsiesFighterState", "FootsiesMove", "FootsiesMoveInfo",
           "FOOTSIES_MOVE_INDEX_TO_MOVE", "FOOTSIES_MOVE_ID_TO_INDEX"]
__version__ = "0.1.0"
