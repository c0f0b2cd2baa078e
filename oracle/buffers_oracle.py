"""TEST INFRASTRUCTURE — numpy restatement of the reference's replay buffers (host arrays, no device).

Restates the *algorithm* of sheeprl/data/buffers.py (ReplayBuffer :20-361, SequentialReplayBuffer :364-526,
EnvIndependentReplayBuffer :529-743): ring write, valid start ranges around the write head, Generator call
sequence, row take, [n_samples,(T,)B,...] layout.  Parity PINNED: tests/golden/buffers_*.npz hold indices and
sampled bytes produced by the executed reference (oracle/make_golden_buffers.py, run in the build container where
/root/reference exists); tests/test_buffers_cpu.py checks this file against them bit for bit.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may import this module.
"""
from __future__ import annotations

import numpy as np


class RingOracle:
    """ReplayBuffer / SequentialReplayBuffer semantics on numpy arrays."""

    def __init__(self, size: int, n_envs: int = 1, obs_keys=("observations",), sequential: bool = False, seed=None):
        self.size, self.n_envs, self.obs_keys, self.sequential = size, n_envs, tuple(obs_keys), sequential
        self.buf = self._buf = {}
        self.pos, self.full = 0, False
        self._rng = np.random.default_rng(seed)

    # buffers.py:145-221
    def add(self, data):
        n = next(iter(data.values())).shape[0]
        nxt = (self.pos + n) % self.size
        if nxt <= self.pos or (n > self.size and not self.full):
            rows = list(range(self.pos, self.size)) + list(range(0, nxt))
        else:
            rows = list(range(self.pos, nxt))
        rows = np.asarray(rows, dtype=np.int64)
        for k, v in data.items():
            v = v[-self.size - nxt:] if n > self.size else v
            if k not in self.buf:
                self.buf[k] = np.empty((self.size, self.n_envs, *v.shape[2:]), dtype=v.dtype)
            self.buf[k][rows] = v
        if self.pos + n >= self.size:
            self.full = True
        self.pos = nxt

    def _valid(self, first_end):
        second_end = self.size if first_end >= 0 else self.size + first_end
        return np.asarray(list(range(0, first_end)) + list(range(self.pos, second_end)), dtype=np.intp)

    # buffers.py:223-288 (uniform) / :395-526 (sequential)
    def sample(self, batch_size, sample_next_obs=False, n_samples=1, sequence_length=1):
        n = batch_size * n_samples
        if not self.sequential:
            if self.full:
                valid = self._valid(self.pos - 1 if sample_next_obs else self.pos)
                t_idx = valid[self._rng.integers(0, len(valid), size=(n,), dtype=np.intp)]
            else:
                t_idx = self._rng.integers(0, self.pos - 1 if sample_next_obs else self.pos, size=(n,), dtype=np.intp)
            e_idx = self._rng.integers(0, self.n_envs, size=(n,), dtype=np.intp)
            out = {}
            for k, v in self.buf.items():
                out[k] = v[t_idx, e_idx].reshape(n_samples, batch_size, *v.shape[2:])
                if sample_next_obs and k in self.obs_keys:
                    out[f"next_{k}"] = v[(t_idx + 1) % self.size, e_idx].reshape(n_samples, batch_size, *v.shape[2:])
            return out, t_idx * self.n_envs + e_idx
        T = sequence_length
        if self.full:
            valid = self._valid(self.pos - T + 1)
            start = valid[self._rng.integers(0, len(valid), size=(n,), dtype=np.intp)]
        else:
            start = self._rng.integers(0, self.pos - T + 1, size=(n,), dtype=np.intp)
        t_idx = (start[:, None] + np.arange(T)[None, :]) % self.size             # [n, T]
        if self.n_envs == 1:
            e_idx = np.zeros((n, 1), dtype=np.intp)
        else:
            e_idx = self._rng.integers(0, self.n_envs, size=(n,), dtype=np.intp)[:, None]
        e_idx = np.broadcast_to(e_idx, t_idx.shape)
        out = {}
        for k, v in self.buf.items():
            g = v[t_idx, e_idx].reshape(n_samples, batch_size, T, *v.shape[2:])
            out[k] = np.swapaxes(g, 1, 2)
            if sample_next_obs:
                g = v[(t_idx + 1) % self.size, e_idx].reshape(n_samples, batch_size, T, *v.shape[2:])
                out[f"next_{k}"] = np.swapaxes(g, 1, 2)
        return out, (t_idx * self.n_envs + e_idx).reshape(-1)


class EnvIndependentOracle:
    """buffers.py:529-743: independent rings (n_envs=1 each), batch split by a bincount of uniform env draws."""

    def __init__(self, size: int, n_envs: int, obs_keys=("observations",), sequential: bool = False, seed=None):
        self._buf = [RingOracle(size, 1, obs_keys, sequential) for _ in range(n_envs)]
        self._rng = np.random.default_rng(seed)
        self.n_envs, self.axis = n_envs, 2 if sequential else 1

    def add(self, data, indices=None):
        indices = range(self.n_envs) if indices is None else indices
        for j, e in enumerate(indices):
            self._buf[e].add({k: v[:, j:j + 1] for k, v in data.items()})

    def sample(self, batch_size, sample_next_obs=False, n_samples=1, **kw):
        per = np.bincount(self._rng.integers(0, self.n_envs, (batch_size,)))
        parts = [r.sample(int(b), sample_next_obs, n_samples, **kw)[0] for r, b in zip(self._buf, per) if b > 0]
        return {k: np.concatenate([p[k] for p in parts], axis=self.axis) for k in parts[0]}
