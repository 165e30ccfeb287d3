"""footsies_gym_b200 -- B200-native batched FOOTSIES simulator behind the FootsiesEnv API.

Importing the package never touches CUDA; constructing a FootsiesEnv does, and fails loudly without a
GPU or without the in-tree library (python -m footsies_gym_b200.build).
"""
from .env import FootsiesEnv, FootsiesGameClosedError
from .state import FootsiesBattleState, FootsiesFighterState, FootsiesState
from .moves import FOOTSIES_MOVE_ID_TO_INDEX, FOOTSIES_MOVE_INDEX_TO_MOVE, FootsiesMove, FootsiesMoveInfo

__all__ = ["FootsiesEnv", "FootsiesGameClosedError", "FootsiesState", "FootsiesBattleState", "FootsiesFighterState", "FootsiesMove", "FootsiesMoveInfo",
           "FOOTSIES_MOVE_INDEX_TO_MOVE", "FOOTSIES_MOVE_ID_TO_INDEX"]
__version__ = "0.2.0"

# footsies_gym/__init__.py:3-7: the reference registers itself with gymnasium on import.  gymnasium is optional here; when
# it is importable the batched env is registered under the same id (it is deterministic given its seed, unlike the
# reference, whose `nondeterministic=True` comes from its asynchronous SEED command).
try:  # pragma: no cover - depends on the environment
    from gymnasium.envs.registration import register as _register, registry as _registry
    if "FootsiesEnv-v0" not in _registry:
        _register(id="FootsiesEnv-v0", entry_point="footsies_gym_b200.env:FootsiesEnv", nondeterministic=False,
                  disable_env_checker=True)        # batched tensors, not a single-env numpy observation
except ImportError:
    pass
