"""Minimal stand-ins for the four gym names the marinenav env uses (gym.Env, spaces.Discrete, spaces.Box,
envs.registration.register) plus make().  Used ONLY when the real `gym` package is not installed
(marinenav_env.py:4,25,33-37 and marinenav_env/__init__.py:1-5 of the reference)."""
import importlib

import numpy as np

_REGISTRY = {}


class Env:
    metadata = {}

    def close(self):
        pass


class Discrete:
    def __init__(self, n):
        self.n = int(n)
        self.shape = ()
        self.dtype = np.int64

    def sample(self):
        return int(np.random.randint(self.n))

    def contains(self, x):
        return 0 <= int(x) < self.n


class Box:
    def __init__(self, low, high, dtype=np.float32, shape=None):
        self.low, self.high, self.dtype = np.asarray(low), np.asarray(high), dtype
        self.shape = self.low.shape if shape is None else shape


class _Spaces:
    Discrete, Box = Discrete, Box


spaces = _Spaces()


def register(id, entry_point, **kwargs):
    _REGISTRY[id] = (entry_point, kwargs)


def make(id, **kwargs):
    """gym.make('marinenav_env:marinenav_env-v0', seed=..., schedule=...) (train_IQN_model.py:96)."""
    if ":" in id:
        module, id = id.split(":", 1)
        importlib.import_module(module)
    entry_point, defaults = _REGISTRY[id]
    mod, cls = entry_point.split(":")
    ctor = getattr(importlib.import_module(mod), cls)
    return ctor(**{**defaults, **kwargs})
