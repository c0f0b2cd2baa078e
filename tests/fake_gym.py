"""TEST INFRASTRUCTURE — the few pieces of `gymnasium` the reference's `main` loops touch (spaces for isinstance checks,
SyncVectorEnv), so that the UNMODIFIED reference `main` can be driven on a host without gymnasium."""
from __future__ import annotations

import types

import numpy as np


class Space:
    pass


class Box(Space):
    def __init__(self, low, high, shape, dtype=np.float32):
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype

    def sample(self):
        return np.random.uniform(self.low, self.high, self.shape).astype(self.dtype)


class Discrete(Space):
    def __init__(self, n):
        self.n, self.shape = int(n), ()

    def sample(self):
        return np.int64(np.random.randint(self.n))


class MultiDiscrete(Space):
    def __init__(self, nvec):
        self.nvec = np.asarray(nvec)
        self.shape = self.nvec.shape

    def sample(self):
        return np.array([np.random.randint(n) for n in self.nvec], dtype=np.int64)


class Dict(Space, dict):
    def __init__(self, spaces):
        dict.__init__(self, spaces)

    @property
    def spaces(self):
        return self


class _VecActionSpace:
    def __init__(self, single, n):
        self.single, self.n = single, n
        self.shape = (n,) + tuple(single.shape)

    def sample(self):
        return np.stack([self.single.sample() for _ in range(self.n)])


class SyncVectorEnv:
    def __init__(self, fns):
        self.envs = [f() for f in fns]
        self.single_action_space = self.envs[0].action_space
        self.single_observation_space = self.envs[0].observation_space
        self.action_space = _VecActionSpace(self.single_action_space, len(self.envs))

    def _stack(self, obs):
        return {k: np.stack([o[k] for o in obs]) for k in obs[0]}

    def reset(self, seed=None):
        return self._stack([e.reset(seed=seed)[0] for e in self.envs]), {}

    def step(self, actions):
        out = [e.step(a) for e, a in zip(self.envs, actions)]
        obs, rew, term, trunc = [o[0] for o in out], [o[1] for o in out], [o[2] for o in out], [o[3] for o in out]
        for i, (t, u) in enumerate(zip(term, trunc)):
            if t or u:                       # autoreset: the loop reads the post-reset observation (the tests do not
                obs[i] = self.envs[i].reset()[0]   # exercise final_observation)
        return self._stack(obs), np.array(rew, dtype=np.float32), np.array(term), np.array(trunc), {}

    def close(self):
        pass


class DummyImageEnv:
    """64x64x3 uint8 observations, 2 discrete actions, episodes of `length` steps (the shape of the reference's
    `discrete_dummy` env, sheeprl/envs/dummy.py:81-93)."""

    def __init__(self, size=64, length=12, seed=0, continuous=False, vector_dim=0):
        obs = {"rgb": Box(0, 255, (3, size, size), np.uint8)}
        if vector_dim:
            obs["state"] = Box(-1, 1, (vector_dim,), np.float32)
        self.observation_space = Dict(obs)
        self.action_space = Box(-1.0, 1.0, (2,), np.float32) if continuous else Discrete(2)
        self.length, self.t = length, 0
        self.rng = np.random.default_rng(seed)
        self.size, self.vector_dim = size, vector_dim

    def _obs(self):
        o = {"rgb": self.rng.integers(0, 256, (3, self.size, self.size), dtype=np.uint8)}
        if self.vector_dim:
            o["state"] = self.rng.uniform(-1, 1, self.vector_dim).astype(np.float32)
        return o

    def reset(self, seed=None, options=None):
        self.t = 0
        return self._obs(), {}

    def step(self, action):
        self.t += 1
        return self._obs(), float(self.rng.normal()), self.t >= self.length, False, {}

    def close(self):
        pass


def module():
    m = types.ModuleType("gymnasium")
    m.spaces = types.SimpleNamespace(Box=Box, Discrete=Discrete, MultiDiscrete=MultiDiscrete, Dict=Dict, Space=Space)
    m.vector = types.SimpleNamespace(SyncVectorEnv=SyncVectorEnv, AsyncVectorEnv=SyncVectorEnv)
    m.Env = object
    return m
