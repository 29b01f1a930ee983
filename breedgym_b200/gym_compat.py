"""Gymnasium surface: the real package when importable, else a minimal shim.

The reference subclasses `gymnasium.Env`, `gymnasium.Wrapper`,
`gymnasium.experimental.vector.{VectorEnv,VectorWrapper}` and registers six ids
(breedgym/__init__.py:4-34).  gymnasium is not installed on the build image or
the GPU box, so this module supplies just the attributes BreedGym relies on
(SURVEY.md App. D): `Env.reset(seed)` -> `np_random` (PCG64 of a SeedSequence),
`Wrapper` attribute forwarding, the spaces used by the envs with faithful
`sample()` semantics, and `register` / `make` with `"module:Id"` resolution.
"""
from __future__ import annotations

import importlib
from typing import Any, Dict, Optional, Sequence as _Seq

import numpy as np

try:  # pragma: no cover - depends on the environment
    import gymnasium as _gym
    from gymnasium import spaces
    from gymnasium.envs.registration import register

    try:
        from gymnasium.experimental.vector import VectorEnv, VectorWrapper  # gymnasium 0.29
    except ImportError:
        from gymnasium.vector import VectorEnv, VectorWrapper  # gymnasium >= 1.0
    Env, Wrapper, make = _gym.Env, _gym.Wrapper, _gym.make
    HAVE_GYMNASIUM = True
except ImportError:
    HAVE_GYMNASIUM = False

    def _np_random(seed=None):
        seq = np.random.SeedSequence(seed)
        return np.random.Generator(np.random.PCG64(seq)), seq.entropy

    class _Space:
        def __init__(self, shape=None, dtype=None, seed=None):
            self._shape = None if shape is None else tuple(shape)
            self.dtype = None if dtype is None else np.dtype(dtype)
            self._np_random = None
            if seed is not None:
                self.seed(seed)

        @property
        def shape(self):
            return self._shape

        @property
        def np_random(self):
            if self._np_random is None:
                self.seed()
            return self._np_random

        def seed(self, seed=None):
            self._np_random, s = _np_random(seed)
            return [s]

        def sample(self):  # pragma: no cover
            raise NotImplementedError

        def contains(self, x) -> bool:  # pragma: no cover
            raise NotImplementedError

        def __contains__(self, x):
            return self.contains(x)

    class Box(_Space):
        def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
            if shape is None:
                shape = np.broadcast(np.asarray(low), np.asarray(high)).shape
            super().__init__(shape, dtype, seed)
            self.low = np.broadcast_to(np.asarray(low), self._shape)
            self.high = np.broadcast_to(np.asarray(high), self._shape)

        def sample(self):
            if self.dtype.kind == "f":
                return self.np_random.uniform(self.low, self.high, size=self._shape).astype(self.dtype)
            if self.dtype.kind == "b":
                return self.np_random.integers(0, 2, size=self._shape).astype(self.dtype)
            return self.np_random.integers(self.low, self.high, size=self._shape, endpoint=True).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self._shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

        def __repr__(self):
            return f"Box({self.low.min()}, {self.high.max()}, {self._shape}, {self.dtype})"

    class Discrete(_Space):
        def __init__(self, n, seed=None, start=0):
            super().__init__((), np.int64, seed)
            self.n, self.start = int(n), int(start)

        def sample(self, mask=None):
            return int(self.start + self.np_random.integers(self.n))

        def contains(self, x):
            try:
                xi = int(x)
            except (TypeError, ValueError):
                return False
            return xi == x and self.start <= xi < self.start + self.n

        def __repr__(self):
            return f"Discrete({self.n}, start={self.start})" if self.start else f"Discrete({self.n})"

    class MultiDiscrete(_Space):
        def __init__(self, nvec, dtype=np.int64, seed=None):
            self.nvec = np.asarray(nvec, dtype=dtype)
            super().__init__(self.nvec.shape, dtype, seed)

        def sample(self, mask=None):
            return (self.np_random.random(self.nvec.shape) * self.nvec).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.nvec.shape and bool(np.all(x >= 0) and np.all(x < self.nvec))

    class Tuple(_Space):
        def __init__(self, spaces, seed=None):
            self.spaces = tuple(spaces)
            super().__init__(None, None, seed)

        def seed(self, seed=None):
            out = super().seed(seed)
            for i, s in enumerate(self.spaces):
                s.seed(int(self._np_random.integers(np.iinfo(np.int32).max)))
            return out

        def sample(self, mask=None):
            return tuple(s.sample() for s in self.spaces)

        def contains(self, x):
            return len(x) == len(self.spaces) and all(s.contains(v) for s, v in zip(self.spaces, x))

        def __len__(self):
            return len(self.spaces)

        def __getitem__(self, i):
            return self.spaces[i]

    class Dict(_Space):
        def __init__(self, spaces=None, seed=None, **kw):
            self.spaces = dict(spaces or {}, **kw)
            super().__init__(None, None, seed)

        def seed(self, seed=None):
            out = super().seed(seed)
            for s in self.spaces.values():
                s.seed(int(self._np_random.integers(np.iinfo(np.int32).max)))
            return out

        def sample(self, mask=None):
            return {k: s.sample() for k, s in self.spaces.items()}

        def contains(self, x):
            return isinstance(x, dict) and x.keys() == self.spaces.keys() and all(
                self.spaces[k].contains(v) for k, v in x.items())

        def __getitem__(self, k):
            return self.spaces[k]

        def keys(self):
            return self.spaces.keys()

    class Sequence(_Space):
        """Variable-length sequences; `sample()` draws length ~ Geometric(0.25)."""

        def __init__(self, space, seed=None, stack=False):
            self.feature_space = space
            self.stack = stack
            super().__init__(None, None, seed)

        def seed(self, seed=None):
            out = super().seed(seed)
            self.feature_space.seed(int(self._np_random.integers(np.iinfo(np.int32).max)))
            return out

        def sample(self, mask=None):
            length = int(self.np_random.geometric(0.25))
            return tuple(self.feature_space.sample() for _ in range(length))

        def contains(self, x):
            return all(self.feature_space.contains(v) for v in x)

    class _Spaces:
        pass

    spaces = _Spaces()
    for _c in (Box, Discrete, MultiDiscrete, Tuple, Dict, Sequence):
        setattr(spaces, _c.__name__, _c)
    spaces.Space = _Space

    class Env:
        metadata: Dict[str, Any] = {"render_modes": []}
        render_mode: Optional[str] = None
        observation_space = None
        action_space = None
        _np_random = None

        @property
        def np_random(self):
            if self._np_random is None:
                self._np_random, _ = _np_random()
            return self._np_random

        @np_random.setter
        def np_random(self, value):
            self._np_random = value

        @property
        def unwrapped(self):
            return self

        def reset(self, *, seed=None, options=None):
            if seed is not None:
                self._np_random, _ = _np_random(seed)

        def step(self, action):  # pragma: no cover
            raise NotImplementedError

        def render(self):
            return None

        def close(self):
            pass

    def _forwarded(name):
        priv = "_own_" + name

        def fget(self):
            own = self.__dict__.get(priv)
            return getattr(self.env, name) if own is None else own

        def fset(self, value):
            self.__dict__[priv] = value

        return property(fget, fset)

    class Wrapper(Env):
        observation_space = _forwarded("observation_space")
        action_space = _forwarded("action_space")
        render_mode = _forwarded("render_mode")

        def __init__(self, env):
            self.env = env

        def __getattr__(self, name):
            if name.startswith("_") or name == "env":
                raise AttributeError(name)
            return getattr(self.env, name)

        @property
        def unwrapped(self):
            return self.env.unwrapped

        @property
        def np_random(self):
            return self.env.np_random

        @np_random.setter
        def np_random(self, value):
            self.env.np_random = value

        def reset(self, *, seed=None, options=None):
            return self.env.reset(seed=seed, options=options)

        def step(self, action):
            return self.env.step(action)

        def render(self):
            return self.env.render()

        def close(self):
            return self.env.close()

    class VectorEnv:
        num_envs: int = 1
        observation_space = None
        action_space = None
        single_observation_space = None
        single_action_space = None
        closed = False

        @property
        def unwrapped(self):
            return self

        def reset(self, *, seed=None, options=None):  # pragma: no cover
            raise NotImplementedError

        def step(self, actions):  # pragma: no cover
            raise NotImplementedError

        def close(self, **kwargs):
            self.closed = True

    class VectorWrapper(VectorEnv):
        num_envs = _forwarded("num_envs")
        observation_space = _forwarded("observation_space")
        action_space = _forwarded("action_space")
        single_observation_space = _forwarded("single_observation_space")
        single_action_space = _forwarded("single_action_space")

        def __init__(self, env):
            self.env = env

        def __getattr__(self, name):
            if name.startswith("_") or name == "env":
                raise AttributeError(name)
            return getattr(self.env, name)

        @property
        def unwrapped(self):
            return self.env.unwrapped

        def reset(self, *, seed=None, options=None):
            return self.env.reset(seed=seed, options=options)

        def step(self, actions):
            return self.env.step(actions)

        def close(self, **kwargs):
            return self.env.close(**kwargs)

    _REGISTRY: Dict[str, str] = {}

    def register(id: str, entry_point: str, **kwargs):
        _REGISTRY[id] = entry_point

    def make(id: str, **kwargs):
        if ":" in id:
            module, id = id.split(":", 1)
            importlib.import_module(module)
        if id not in _REGISTRY:
            raise KeyError(f"No registered env with id: {id}")
        mod, attr = _REGISTRY[id].split(":")
        return getattr(importlib.import_module(mod), attr)(**kwargs)
