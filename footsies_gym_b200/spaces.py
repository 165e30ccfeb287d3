"""Observation / action space descriptions.

The reference builds gymnasium spaces (footsies.py:157-171).  gymnasium is optional here: when it is
importable the real classes are used, otherwise small duck-typed stand-ins with the same attributes
(`nvec`, `low`, `high`, `shape`, `n`, `spaces`, `contains`, `sample`) take their place.
"""
import numpy as np

try:  # pragma: no cover - depends on the environment
    from gymnasium import spaces as _gs
    HAVE_GYMNASIUM = True
except Exception:  # noqa: BLE001
    _gs = None
    HAVE_GYMNASIUM = False


class _Space:
    def __init__(self, shape, dtype):
        self.shape = tuple(shape)
        self.dtype = np.dtype(dtype)
        self._rng = np.random.default_rng()

    def seed(self, seed=None):
        self._rng = np.random.default_rng(seed)


class MultiDiscrete(_Space):
    def __init__(self, nvec):
        self.nvec = np.asarray(nvec, dtype=np.int64)
        super().__init__(self.nvec.shape, np.int64)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all((x >= 0) & (x < self.nvec)))

    def sample(self):
        return (self._rng.random(self.shape) * self.nvec).astype(np.int64)

    def __repr__(self):
        return f"MultiDiscrete({self.nvec.tolist()})"


class Box(_Space):
    def __init__(self, low, high, shape, dtype=np.float32):
        super().__init__(shape, dtype)
        self.low = np.full(shape, low, dtype=dtype)
        self.high = np.full(shape, high, dtype=dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all((x >= self.low) & (x <= self.high)))

    def sample(self):
        return (self.low + self._rng.random(self.shape) * (self.high - self.low)).astype(self.dtype)

    def __repr__(self):
        return f"Box({self.low.flat[0]}, {self.high.flat[0]}, {self.shape})"


class MultiBinary(_Space):
    def __init__(self, n):
        self.n = n
        super().__init__((n,), np.int8)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all((x == 0) | (x == 1)))

    def sample(self):
        return self._rng.integers(0, 2, size=self.shape).astype(np.int8)

    def __repr__(self):
        return f"MultiBinary({self.n})"


class Discrete(_Space):
    def __init__(self, n):
        self.n = int(n)
        super().__init__((), np.int64)

    def contains(self, x):
        return 0 <= int(x) < self.n

    def sample(self):
        return int(self._rng.integers(0, self.n))

    def __repr__(self):
        return f"Discrete({self.n})"


class Dict(_Space):
    def __init__(self, spaces):
        self.spaces = dict(spaces)
        super().__init__((), np.float32)

    def __getitem__(self, k):
        return self.spaces[k]

    def keys(self):
        return self.spaces.keys()

    def contains(self, x):
        return set(x.keys()) == set(self.spaces.keys()) and all(self.spaces[k].contains(v) for k, v in x.items())

    def sample(self):
        return {k: s.sample() for k, s in self.spaces.items()}

    def __repr__(self):
        return "Dict(" + ", ".join(f"{k!r}: {v!r}" for k, v in self.spaces.items()) + ")"


if HAVE_GYMNASIUM:  # pragma: no cover
    MultiDiscrete, Box, MultiBinary, Discrete, Dict = (_gs.MultiDiscrete, _gs.Box, _gs.MultiBinary,
                                                       _gs.Discrete, _gs.Dict)


def footsies_observation_space(num_moves=15, max_move_duration=55.0):
    """footsies.py:153-168: WIN and DEAD are not 'relevant' moves; move_frame high = longest relevant move."""
    return Dict({
        "guard": MultiDiscrete([4, 4]),
        "move": MultiDiscrete([num_moves, num_moves]),
        "move_frame": Box(low=0.0, high=float(max_move_duration), shape=(2,)),
        "position": Box(low=-4.6, high=4.6, shape=(2,)),
    })


def footsies_action_space():
    """footsies.py:171: left, right, attack."""
    return MultiBinary(3)
