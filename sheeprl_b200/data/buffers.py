"""Device-resident replay storage behind the reference's buffer API.

Mirrors (behaviour, names, argument meaning, errors):
  * ``ReplayBuffer``                sheeprl/data/buffers.py:20-361   (SAC / PPO storage, uniform rows)
  * ``SequentialReplayBuffer``      sheeprl/data/buffers.py:364-526  (Dreamer sequences)
  * ``EnvIndependentReplayBuffer``  sheeprl/data/buffers.py:529-743  (one ring per environment)
  * ``get_tensor``                  sheeprl/data/buffers.py:1158-1180

Design (B200-first, not a translation):
  * every key lives in ONE HBM allocation ``[n_slabs * buffer_size * n_envs, row_bytes]`` in the dtype of the first
    ``add`` (as the reference does, buffers.py:203-215); 1 M Atari-sized rows = 12.3 GB of the 180 GB;
  * the *index plan* stays on the host and draws from the same ``numpy.random.Generator`` with the same call
    sequence, arguments and dtypes as the reference, so sampled indices are bit-identical; the list-of-ranges the
    reference materialises (O(buffer_size) Python objects per sample call) is replaced by closed-form arithmetic on
    the drawn positions;
  * bytes only move on the device: one ``b200rl_replay_gather`` launch per key emits ``[n_samples, T, B, ...]``
    directly (the reference's np.take + reshape + swapaxes + concatenate + H2D), one ``b200rl_replay_scatter`` per
    key does the ring write.  ``EnvIndependentReplayBuffer`` shares one allocation between its per-env rings so
    that a whole Dreamer batch is still a single launch per key.

There is no CPU data path here: ``ops`` must provide the kernels (``sheeprl_b200.lib.CudaOps``; the tests on a
GPU-less host pass the torch test double from ``oracle/ops_emul.py``).
"""
from __future__ import annotations

import json
import os
from typing import Dict, Optional, Sequence, Type

import numpy as np
import torch

_MEMMAP_MODES = ("r+", "w+", "c", "copyonwrite", "readwrite", "write")


def _restore_device(name: str) -> torch.device:
    """device of an unpickled buffer: where it was, or the host when that accelerator is absent (the kernels then
    refuse to run, as everywhere else)"""
    dev = torch.device(name)
    if dev.type == "cuda" and not torch.cuda.is_available():
        return torch.device("cpu")
    return dev


def _clone_rng(rng: np.random.Generator) -> np.random.Generator:
    out = np.random.Generator(type(rng.bit_generator)())
    out.bit_generator.state = rng.bit_generator.state
    return out


# Process-wide defaults for buffers built WITHOUT a `device` / `ops` argument — which is how the reference's own `main`
# constructs them (dreamer_v3.py:474-480: it has no such notion).  The B200 entry points set the device to `fabric.device`
# before delegating; tests on a GPU-less host set `ops` to the torch test double.
DEFAULTS = {"device": "cuda", "ops": None}


def _default_ops():
    if DEFAULTS["ops"] is not None:
        return DEFAULTS["ops"]
    from sheeprl_b200.lib import CudaOps  # raises B200RLError when the extension / a B200 is missing

    return CudaOps(DEFAULTS["device"]) if DEFAULTS["device"] != "cuda" else CudaOps()


def get_tensor(array, dtype: Optional[torch.dtype] = None, clone: bool = False, device="cpu", from_numpy: bool = False):
    """numpy / tensor -> tensor on ``device`` (buffers.py:1158-1180).  ``from_numpy`` is accepted for signature
    compatibility; both routes share memory with the source array until the device copy."""
    if isinstance(array, torch.Tensor):
        t = array.clone() if clone else array
    else:
        a = np.ascontiguousarray(array)
        t = torch.from_numpy(a.copy() if clone else a)
    return t.to(device=device, dtype=t.dtype if dtype is None else dtype)


# ---------------------------------------------------------------------------------------------------------
# index plans (host, integer, bit-exact with the reference's Generator call sequence)
# ---------------------------------------------------------------------------------------------------------
def _draw_outside_window(rng: np.random.Generator, n: int, pos: int, size: int, first_range_end: int) -> np.ndarray:
    """Draw ``n`` ring positions from ``[0, first_range_end) U [pos, second_range_end)`` exactly as
    ``valid_idxes[rng.integers(0, len(valid_idxes), ...)]`` does (buffers.py:246-254, 439-453) without building
    ``valid_idxes``: positions below the first range map to themselves, the rest are shifted to start at ``pos``."""
    second_range_end = size if first_range_end >= 0 else size + first_range_end
    n_first = max(first_range_end, 0)
    n_valid = n_first + max(second_range_end - pos, 0)
    r = rng.integers(0, n_valid, size=(n,), dtype=np.intp)
    return np.where(r < n_first, r, r - n_first + pos).astype(np.intp, copy=False)


def ring_write_rows(pos: int, data_len: int, size: int, full: bool):
    """Destination ring rows of ``add`` (buffers.py:186-194): returns (rows, n_tail_rows_of_data_used, next_pos).
    ``rows`` may contain duplicates when ``data_len > size``; later entries win, as numpy fancy assignment does."""
    next_pos = (pos + data_len) % size
    if next_pos <= pos or (data_len > size and not full):
        rows = np.concatenate([np.arange(pos, size, dtype=np.int64), np.arange(0, next_pos, dtype=np.int64)])
    else:
        rows = np.arange(pos, next_pos, dtype=np.int64)
    n_used = min(data_len, size + next_pos) if data_len > size else data_len
    return rows, n_used, next_pos


class ReplayBuffer:
    """Uniform-row replay ring ``[buffer_size, n_envs, ...]`` in HBM (reference: buffers.py:20-361)."""

    batch_axis: int = 1

    def __init__(self, buffer_size: int, n_envs: int = 1, obs_keys: Sequence[str] = ("observations",),
                 memmap: bool = False, memmap_dir=None, memmap_mode: str = "r+", device=None, ops=None, **kwargs):
        device = DEFAULTS["device"] if device is None else device
        if buffer_size <= 0:
            raise ValueError(f"The buffer size must be greater than zero, got: {buffer_size}")
        if n_envs <= 0:
            raise ValueError(f"The number of environments must be greater than zero, got: {n_envs}")
        if memmap:
            # storage is HBM-resident; the flags are validated like the reference (buffers.py:61-75) and otherwise
            # unused (disk spill / restore of the ring is SURVEY §8f rank 2)
            if memmap_mode not in _MEMMAP_MODES:
                raise ValueError('Accepted values for memmap_mode are "r+", "readwrite", "w+", "write", "c" or '
                                 '"copyonwrite".')
            if memmap_dir is None:
                raise ValueError("The buffer is set to be memory-mapped but the 'memmap_dir' attribute is None. "
                                 "Set the 'memmap_dir' to a known directory.")
        self._buffer_size = int(buffer_size)
        self._n_envs = int(n_envs)
        self._obs_keys = tuple(obs_keys)
        self._memmap, self._memmap_dir, self._memmap_mode = memmap, memmap_dir, memmap_mode
        self._device = torch.device(device)
        self._ops = ops
        self._buf: Dict[str, torch.Tensor] = {}
        self._pos = 0
        self._full = False
        self._rng: np.random.Generator = np.random.default_rng()
        # slab sharing (EnvIndependentReplayBuffer): the allocator and this ring's first row in the shared storage
        self._owner = None
        self._slab = 0

    # ------------------------------------------------------------------ properties (reference names)
    @property
    def buffer(self) -> Dict[str, torch.Tensor]:
        return self._buf

    @property
    def buffer_size(self) -> int:
        return self._buffer_size

    @property
    def full(self) -> bool:
        return self._full

    @property
    def n_envs(self) -> int:
        return self._n_envs

    @property
    def empty(self) -> bool:
        return self._buf is None or len(self._buf) == 0

    @property
    def is_memmap(self) -> bool:
        return self._memmap

    def __len__(self) -> int:
        return self._buffer_size

    @property
    def ops(self):
        if self._ops is None:
            self._ops = _default_ops()
        return self._ops

    # ------------------------------------------------------------------ storage
    def _rows_per_ring(self) -> int:
        return self._buffer_size * self._n_envs

    def _allocate(self, key: str, trailing: tuple, dtype: torch.dtype) -> None:
        if self._owner is not None:
            self._buf[key] = self._owner._slab_view(key, self._slab, trailing, dtype)
        else:
            self._buf[key] = torch.empty((self._buffer_size, self._n_envs, *trailing), dtype=dtype, device=self._device)

    def _storage_rows(self, key: str):
        """(flat storage [rows, ...] shared by all slabs, row offset of this ring)"""
        if self._owner is not None:
            return self._owner._flat(key), self._slab * self._rows_per_ring()
        v = self._buf[key]
        return v.view(self._rows_per_ring(), *v.shape[2:]), 0

    def to_tensor(self, dtype=None, clone: bool = False, device="cpu", from_numpy: bool = False):
        return {k: get_tensor(v, dtype=dtype, clone=clone, device=device) for k, v in self._buf.items()}

    # ------------------------------------------------------------------ add
    def add(self, data, validate_args: bool = False) -> None:
        """Ring write of ``[sequence_length, n_envs, ...]`` arrays; the oldest rows are overwritten
        (buffers.py:145-221)."""
        if isinstance(data, ReplayBuffer):
            data = data.buffer
        if validate_args:
            if not isinstance(data, dict):
                raise ValueError(f"'data' must be a dictionary containing Numpy arrays, but 'data' is of type '{type(data)}'")
            for k, v in data.items():
                if not isinstance(v, (np.ndarray, torch.Tensor)):
                    raise ValueError("'data' must be a dictionary containing Numpy arrays. Found key "
                                     f"'{k}' containing a value of type '{type(v)}'")
            first_key, first_shape = None, None
            for k, v in data.items():
                if len(v.shape) < 2:
                    raise RuntimeError("'data' must have at least 2 dimensions: [sequence_length, n_envs, ...]. "
                                       f"Shape of '{k}' is {tuple(v.shape)}")
                if first_key is None:
                    first_key, first_shape = k, tuple(v.shape[:2])
                elif tuple(v.shape[:2]) != first_shape:
                    raise RuntimeError("Every array in 'data' must be congruent in the first 2 dimensions: "
                                       f"found key '{first_key}' with shape '{first_shape}' "
                                       f"and '{k}' with shape '{tuple(v.shape[:2])}'")
                    # (the reference reports the previous key; the condition is the same)
        data_len = next(iter(data.values())).shape[0]
        rows, n_used, next_pos = ring_write_rows(self._pos, data_len, self._buffer_size, self._full)
        if n_used != len(rows):
            # same failure the reference hits in `buffer[idxes] = data_to_store` (buffers.py:196-216)
            raise ValueError(f"shape mismatch: value array of shape ({n_used},...) could not be broadcast to indexing "
                             f"result of shape ({len(rows)},...)")
        # later duplicates win (numpy fancy-assignment order); drop the earlier ones so that the scatter has no race
        keep = np.ones(len(rows), dtype=bool)
        if data_len > self._buffer_size:
            _, last = np.unique(rows[::-1], return_index=True)
            keep[:] = False
            keep[len(rows) - 1 - last] = True
        time_rows = rows[keep]
        # ring row r, env e -> storage row r * n_envs + e
        dst = (time_rows[:, None] * self._n_envs + np.arange(self._n_envs, dtype=np.int64)[None, :]).reshape(-1)
        first_add = self.empty
        for k, v in data.items():
            t = v if isinstance(v, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(v))
            t = t[data_len - n_used:]
            if not bool(keep.all()):
                t = t[torch.from_numpy(np.nonzero(keep)[0])]
            if k not in self._buf:
                if not first_add:
                    raise KeyError(f"key '{k}' was not part of the first add()")
                self._allocate(k, tuple(t.shape[2:]), t.dtype)
            store, off = self._storage_rows(k)
            src = t.to(device=store.device, dtype=store.dtype, non_blocking=True).contiguous()
            src = src.view(-1, *store.shape[1:])
            dst_dev = torch.from_numpy(dst + off).to(store.device, non_blocking=True)
            self.ops.replay_scatter(src, dst_dev, store)
        if self._pos + data_len >= self._buffer_size:
            self._full = True
        self._pos = next_pos

    # ------------------------------------------------------------------ sample
    def _plan(self, batch_size: int, sample_next_obs: bool, n_samples: int):
        """Ring/env indices of one ``sample`` call — the reference's RNG call sequence (buffers.py:223-274)."""
        if batch_size <= 0 or n_samples <= 0:
            raise ValueError(f"'batch_size' ({batch_size}) and 'n_samples' ({n_samples}) must be both greater than 0")
        if not self._full and self._pos == 0:
            raise ValueError("No sample has been added to the buffer. Please add at least one sample calling 'self.add()'")
        n = batch_size * n_samples
        if self._full:
            first_range_end = self._pos - 1 if sample_next_obs else self._pos
            batch_idxes = _draw_outside_window(self._rng, n, self._pos, self._buffer_size, first_range_end)
        else:
            max_pos_to_sample = self._pos - 1 if sample_next_obs else self._pos
            if max_pos_to_sample == 0:
                raise RuntimeError("You want to sample the next observations, but one sample has been added to the "
                                   "buffer. Make sure that at least two samples are added.")
            batch_idxes = self._rng.integers(0, max_pos_to_sample, size=(n,), dtype=np.intp)
        if self.empty:
            raise RuntimeError("The buffer has not been initialized. Try to add some data first.")
        env_idxes = self._rng.integers(0, self._n_envs, size=(n,), dtype=np.intp)
        rows = batch_idxes * self._n_envs + env_idxes
        next_rows = ((batch_idxes + 1) % self._buffer_size) * self._n_envs + env_idxes if sample_next_obs else None
        return rows.astype(np.int64, copy=False), None if next_rows is None else next_rows.astype(np.int64, copy=False)

    def _gather(self, key: str, rows_dev: torch.Tensor, n_samples: int, batch: int, seq_len: int) -> torch.Tensor:
        store, _ = self._storage_rows(key)
        shape = (n_samples, batch) if seq_len is None else (n_samples, seq_len, batch)
        out = torch.empty((*shape, *store.shape[1:]), dtype=store.dtype, device=store.device)
        self.ops.replay_gather(store, rows_dev, out, n_samples, batch, 1 if seq_len is None else seq_len)
        return out

    def _to_dev(self, rows: np.ndarray) -> torch.Tensor:
        store, off = self._storage_rows(next(iter(self._buf)))
        return torch.from_numpy(rows + off).to(store.device, non_blocking=True)

    def sample_tensors(self, batch_size: int, clone: bool = False, sample_next_obs: bool = False, dtype=None,
                       device=None, from_numpy: bool = False, **kwargs) -> Dict[str, torch.Tensor]:
        """``[n_samples, batch_size, ...]`` tensors gathered on the device (buffers.py:290-327).  ``clone`` and
        ``from_numpy`` are no-ops: the gather always writes fresh tensors."""
        n_samples = kwargs.pop("n_samples", 1)
        rows, next_rows = self._plan(batch_size, sample_next_obs, n_samples)
        rows_d = self._to_dev(rows)
        next_d = self._to_dev(next_rows) if next_rows is not None else None
        out = {}
        for k in self._buf:
            out[k] = self._gather(k, rows_d, n_samples, batch_size, None)
            if sample_next_obs and k in self._obs_keys:
                out[f"next_{k}"] = self._gather(k, next_d, n_samples, batch_size, None)
        return _finish(out, dtype, device)

    def sample(self, batch_size: int, sample_next_obs: bool = False, clone: bool = False, n_samples: int = 1,
               **kwargs) -> Dict[str, np.ndarray]:
        """numpy view of :meth:`sample_tensors` for callers that want host arrays (compatibility path: it pays a
        device->host copy; the training loop should call ``sample_tensors``)."""
        t = self.sample_tensors(batch_size, sample_next_obs=sample_next_obs, n_samples=n_samples, **kwargs)
        return {k: v.cpu().numpy() for k, v in t.items()}

    # ------------------------------------------------------------------ disk spill / restore
    def to_memmap(self, directory) -> None:
        """Writes the ring to `<directory>/<key>.memmap` — raw `[buffer_size, n_envs, ...]` arrays in the stored dtype,
        the very files the reference's `MemmapArray` keeps when `buffer.memmap=True` (sheeprl/utils/memmap.py:22-85,
        buffers.py:203-210) — plus `meta.json` with the write head, so a device ring can be checkpointed / resumed or
        handed to the reference's host buffers."""
        os.makedirs(directory, exist_ok=True)
        meta = {"buffer_size": self._buffer_size, "n_envs": self._n_envs, "pos": self._pos, "full": self._full, "keys": {}}
        for k, v in self._buf.items():
            a = v.detach().cpu().numpy()
            mm = np.memmap(os.path.join(directory, f"{k}.memmap"), dtype=a.dtype, mode="w+", shape=a.shape)
            mm[:] = a
            mm.flush()
            meta["keys"][k] = {"dtype": str(a.dtype), "shape": list(a.shape)}
        with open(os.path.join(directory, "meta.json"), "w") as f:
            json.dump(meta, f)

    def load_memmap(self, directory) -> None:
        """Inverse of :meth:`to_memmap` (also reads a directory written by the reference when given its key dtypes /
        shapes through `meta.json`)."""
        with open(os.path.join(directory, "meta.json")) as f:
            meta = json.load(f)
        if meta["buffer_size"] != self._buffer_size or meta["n_envs"] != self._n_envs:
            raise ValueError("memmap directory was written by a buffer of a different size")
        for k, info in meta["keys"].items():
            a = np.memmap(os.path.join(directory, f"{k}.memmap"), dtype=np.dtype(info["dtype"]), mode="r",
                          shape=tuple(info["shape"]))
            t = torch.from_numpy(np.ascontiguousarray(a))
            if k not in self._buf:
                self._allocate(k, tuple(t.shape[2:]), t.dtype)
            self._buf[k].copy_(t)
        self._pos, self._full = int(meta["pos"]), bool(meta["full"])

    # ------------------------------------------------------------------ checkpoints
    # The reference checkpoints the buffer OBJECT (`state["rb"] = replay_buffer`, utils/callback.py:37-41), so the
    # device ring pickles as host tensors + write head + Generator and comes back on its device.
    def __getstate__(self):
        st = self.__dict__.copy()
        st["_ops"] = None
        st["_device"] = str(self._device)
        if self._owner is not None:                       # a ring of an EnvIndependentReplayBuffer: the owner holds the bytes
            st["_owner"], st["_buf"] = None, {k: None for k in self._buf}
        else:
            st["_buf"] = {k: v.detach().cpu() for k, v in self._buf.items()}
        return st

    def __setstate__(self, st):
        self.__dict__.update(st)
        self._device = _restore_device(st["_device"])
        self._buf = {k: (v if v is None else v.to(self._device)) for k, v in st["_buf"].items()}

    @classmethod
    def from_reference(cls, ref, device="cuda", ops=None):
        """Adopts a buffer object of the reference (`sheeprl.data.buffers.ReplayBuffer` / `SequentialReplayBuffer`, e.g.
        `state["rb"]` of one of its checkpoints): storage (numpy or `MemmapArray`), write head, fullness, Generator."""
        rb = cls(ref._buffer_size, ref._n_envs, obs_keys=tuple(ref._obs_keys), device=device, ops=ops)
        rb._adopt(ref)
        return rb

    def _adopt(self, ref) -> None:
        for k, v in ref._buf.items():
            t = torch.from_numpy(np.ascontiguousarray(np.asarray(v)))
            self._allocate(k, tuple(t.shape[2:]), t.dtype)
            self._buf[k].copy_(t)
        self._pos, self._full = int(ref._pos), bool(ref._full)
        self._rng = _clone_rng(ref._rng)

    def to_reference(self, memmap: bool = False, memmap_dir=None):
        """The same buffer as an object of the reference's class of the same name (needs `sheeprl` importable), for a
        checkpoint the reference can resume from (`dreamer_v3.py:486-520`)."""
        import importlib

        ref_cls = getattr(importlib.import_module("sheeprl.data.buffers"), type(self).__name__)
        ref = ref_cls(self._buffer_size, self._n_envs, obs_keys=self._obs_keys, memmap=memmap, memmap_dir=memmap_dir)
        self._export(ref)
        return ref

    def _export(self, ref) -> None:
        if self._buf:
            ref.add({k: v.detach().cpu().numpy() for k, v in self._buf.items()})     # allocates (memmap or numpy) + fills
        ref._pos, ref._full = self._pos, self._full
        ref._rng = _clone_rng(self._rng)

    # ------------------------------------------------------------------ item access (reference: buffers.py:329-361)
    def __getitem__(self, key: str) -> torch.Tensor:
        if not isinstance(key, str):
            raise TypeError("'key' must be a string")
        if self.empty:
            raise RuntimeError("The buffer has not been initialized. Try to add some data first.")
        return self._buf.get(key)

    def __setitem__(self, key: str, value) -> None:
        if not isinstance(value, (np.ndarray, torch.Tensor)):
            raise ValueError(f"The value to be set must be an instance of 'np.ndarray' or 'torch.Tensor', got {type(value)}")
        if self.empty:
            raise RuntimeError("The buffer has not been initialized. Try to add some data first.")
        if tuple(value.shape[:2]) != (self._buffer_size, self._n_envs):
            raise RuntimeError("'value' must have at least two dimensions of dimension [buffer_size, n_envs, ...]. "
                               f"Shape of 'value' is {tuple(value.shape)}")
        t = value if isinstance(value, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(value))
        if key not in self._buf:
            self._allocate(key, tuple(t.shape[2:]), t.dtype)
        self._buf[key].copy_(t)


def _finish(out: Dict[str, torch.Tensor], dtype, device) -> Dict[str, torch.Tensor]:
    if dtype is not None or device is not None:
        out = {k: v.to(device=v.device if device is None else device, dtype=v.dtype if dtype is None else dtype)
               for k, v in out.items()}
    return out


class SequentialReplayBuffer(ReplayBuffer):
    """Sequences of consecutive ring rows, ignoring episode boundaries (reference: buffers.py:364-526)."""

    batch_axis: int = 2

    def _plan(self, batch_size: int, sample_next_obs: bool, n_samples: int, sequence_length: int = 1):
        batch_dim = batch_size * n_samples
        if batch_size <= 0 or n_samples <= 0:
            raise ValueError(f"'batch_size' ({batch_size}) and 'n_samples' ({n_samples}) must be both greater than 0")
        if not self._full and self._pos == 0:
            raise ValueError("No sample has been added to the buffer. Please add at least one sample calling 'self.add()'")
        if not self._full and self._pos - sequence_length + 1 < 1:
            raise ValueError(f"Cannot sample a sequence of length {sequence_length}. Data added so far: {self._pos}")
        if self._full and sequence_length > len(self):
            raise ValueError(f"The sequence length ({sequence_length}) is greater than the buffer size ({len(self)})")
        if self._full:
            # never start inside (pos - sequence_length, pos): such a window would straddle the write head
            start = _draw_outside_window(self._rng, batch_dim, self._pos, self._buffer_size,
                                         self._pos - sequence_length + 1)
        else:
            start = self._rng.integers(0, self._pos - sequence_length + 1, size=(batch_dim,), dtype=np.intp)
        ring = (start.reshape(-1, 1) + np.arange(sequence_length, dtype=np.intp).reshape(1, -1)) % self._buffer_size
        if self._n_envs == 1:
            env = np.zeros((batch_dim, 1), dtype=np.intp)
        else:
            env = self._rng.integers(0, self._n_envs, size=(batch_dim,), dtype=np.intp).reshape(-1, 1)
        rows = (ring * self._n_envs + env).reshape(-1)                       # (sample, batch, time) order
        next_rows = (((ring + 1) % self._buffer_size) * self._n_envs + env).reshape(-1) if sample_next_obs else None
        return rows.astype(np.int64, copy=False), None if next_rows is None else next_rows.astype(np.int64, copy=False)

    def sample_tensors(self, batch_size: int, clone: bool = False, sample_next_obs: bool = False, dtype=None,
                       device=None, from_numpy: bool = False, **kwargs) -> Dict[str, torch.Tensor]:
        """``[n_samples, sequence_length, batch_size, ...]`` (buffers.py:395-526)."""
        n_samples = kwargs.pop("n_samples", 1)
        sequence_length = kwargs.pop("sequence_length", 1)
        rows, next_rows = self._plan(batch_size, sample_next_obs, n_samples, sequence_length)
        if self.empty:
            raise RuntimeError("The buffer has not been initialized. Try to add some data first.")
        rows_d = self._to_dev(rows)
        next_d = self._to_dev(next_rows) if next_rows is not None else None
        out = {}
        for k in self._buf:
            out[k] = self._gather(k, rows_d, n_samples, batch_size, sequence_length)
            if sample_next_obs:                                                # every key, as the reference does
                out[f"next_{k}"] = self._gather(k, next_d, n_samples, batch_size, sequence_length)
        return _finish(out, dtype, device)

    def sample(self, batch_size: int, sample_next_obs: bool = False, clone: bool = False, n_samples: int = 1,
               sequence_length: int = 1, **kwargs) -> Dict[str, np.ndarray]:
        t = self.sample_tensors(batch_size, sample_next_obs=sample_next_obs, n_samples=n_samples,
                                sequence_length=sequence_length, **kwargs)
        return {k: v.cpu().numpy() for k, v in t.items()}


class EnvIndependentReplayBuffer:
    """One ring per environment (each with its own write head and Generator), sampled together
    (reference: buffers.py:529-743).  All rings of a key share one HBM allocation ``[n_envs, buffer_size, 1, ...]`` so
    a batch drawn from several environments is still ONE gather launch per key, written straight into the
    concatenated ``[n_samples, (T,) B, ...]`` layout."""

    def __init__(self, buffer_size: int, n_envs: int = 1, obs_keys: Sequence[str] = ("observations",),
                 memmap: bool = False, memmap_dir=None, memmap_mode: str = "r+",
                 buffer_cls: Type[ReplayBuffer] = ReplayBuffer, device=None, ops=None, **kwargs):
        device = DEFAULTS["device"] if device is None else device
        if buffer_size <= 0:
            raise ValueError(f"The buffer size must be greater than zero, got: {buffer_size}")
        if n_envs <= 0:
            raise ValueError(f"The number of environments must be greater than zero, got: {n_envs}")
        if memmap:
            if memmap_mode not in _MEMMAP_MODES:
                raise ValueError('Accepted values for memmap_mode are "r+", "readwrite", "w+", "write", "c" or '
                                 '"copyonwrite".')
            if memmap_dir is None:
                raise ValueError("The buffer is set to be memory-mapped but the 'memmap_dir' attribute is None. "
                                 "Set the 'memmap_dir' to a known directory.")
        self._device = torch.device(device)
        self._ops = ops
        self._store: Dict[str, torch.Tensor] = {}
        self._buf = []
        for i in range(n_envs):
            b = buffer_cls(buffer_size=buffer_size, n_envs=1, obs_keys=obs_keys, memmap=memmap,
                           memmap_dir=memmap_dir, memmap_mode=memmap_mode, device=device, ops=ops, **kwargs)
            b._owner, b._slab = self, i
            self._buf.append(b)
        self._buffer_size = int(buffer_size)
        self._n_envs = int(n_envs)
        self._rng: np.random.Generator = np.random.default_rng()
        self._concat_along_axis = buffer_cls.batch_axis

    # shared-slab allocator used by the per-env rings
    def _slab_view(self, key: str, slab: int, trailing: tuple, dtype: torch.dtype) -> torch.Tensor:
        if key not in self._store:
            self._store[key] = torch.empty((self._n_envs, self._buffer_size, 1, *trailing), dtype=dtype, device=self._device)
        s = self._store[key]
        if tuple(s.shape[3:]) != tuple(trailing) or s.dtype != dtype:
            raise RuntimeError(f"key '{key}': every environment must store the same row shape and dtype")
        return s[slab]

    def _flat(self, key: str) -> torch.Tensor:
        s = self._store[key]
        return s.view(self._n_envs * self._buffer_size, *s.shape[3:])

    @property
    def ops(self):
        if self._ops is None:
            self._ops = _default_ops()
            for b in self._buf:
                b._ops = self._ops
        return self._ops

    @property
    def buffer(self):
        return tuple(self._buf)

    @property
    def buffer_size(self) -> int:
        return self._buffer_size

    @property
    def full(self):
        return tuple(b.full for b in self._buf)

    @property
    def n_envs(self) -> int:
        return self._n_envs

    @property
    def empty(self):
        return tuple(b.empty for b in self._buf)

    @property
    def is_memmap(self):
        return tuple(b.is_memmap for b in self._buf)

    def __len__(self) -> int:
        return self._buffer_size

    def add(self, data, indices: Optional[Sequence[int]] = None, validate_args: bool = False) -> None:
        """``data[:, j]`` goes to ring ``indices[j]`` (buffers.py:636-654)."""
        if isinstance(data, ReplayBuffer):
            data = data.buffer
        if indices is None:
            indices = tuple(range(self._n_envs))
        elif len(indices) != next(iter(data.values())).shape[1]:
            raise ValueError(f"The length of 'indices' ({len(indices)}) must be equal to the second dimension of the "
                             f"arrays in 'data' ({next(iter(data.values())).shape[1]})")
        self.ops  # bind the kernels to every ring
        for j, env_idx in enumerate(indices):
            self._buf[env_idx].add({k: v[:, j:j + 1] for k, v in data.items()}, validate_args=validate_args)

    # ------------------------------------------------------------------ checkpoints (see ReplayBuffer.__getstate__)
    def __getstate__(self):
        st = self.__dict__.copy()
        st["_ops"] = None
        st["_device"] = str(self._device)
        st["_store"] = {k: v.detach().cpu() for k, v in self._store.items()}
        return st

    def __setstate__(self, st):
        self.__dict__.update(st)
        self._device = _restore_device(st["_device"])
        self._store = {k: v.to(self._device) for k, v in st["_store"].items()}
        for i, b in enumerate(self._buf):
            b._owner, b._slab, b._device = self, i, self._device
            b._buf = {k: self._store[k][i] for k in b._buf}

    @classmethod
    def from_reference(cls, ref, device="cuda", ops=None):
        """Adopts a `sheeprl.data.buffers.EnvIndependentReplayBuffer` (e.g. `state["rb"]` of a reference checkpoint)."""
        sequential = ref._concat_along_axis == 2
        rb = cls(ref._buffer_size, ref._n_envs, obs_keys=tuple(ref._buf[0]._obs_keys),
                 buffer_cls=SequentialReplayBuffer if sequential else ReplayBuffer, device=device, ops=ops)
        for mine, theirs in zip(rb._buf, ref._buf):
            mine._adopt(theirs)
        rb._rng = _clone_rng(ref._rng)
        return rb

    def to_reference(self, memmap: bool = False, memmap_dir=None):
        import importlib

        mod = importlib.import_module("sheeprl.data.buffers")
        ref = mod.EnvIndependentReplayBuffer(self._buffer_size, self._n_envs, obs_keys=self._buf[0]._obs_keys, memmap=memmap,
                                             memmap_dir=memmap_dir, buffer_cls=getattr(mod, type(self._buf[0]).__name__))
        for mine, theirs in zip(self._buf, ref._buf):
            mine._export(theirs)
        ref._rng = _clone_rng(self._rng)
        return ref

    def to_memmap(self, directory) -> None:
        """`<directory>/env_<i>/<key>.memmap`, the reference's layout (buffers.py:581)"""
        for i, b in enumerate(self._buf):
            b.to_memmap(os.path.join(directory, f"env_{i}"))

    def load_memmap(self, directory) -> None:
        for i, b in enumerate(self._buf):
            b.load_memmap(os.path.join(directory, f"env_{i}"))

    def sample_tensors(self, batch_size: int, sample_next_obs: bool = False, clone: bool = False, n_samples: int = 1,
                       dtype=None, device=None, from_numpy: bool = False, **kwargs) -> Dict[str, torch.Tensor]:
        if batch_size <= 0 or n_samples <= 0:
            raise ValueError(f"'batch_size' ({batch_size}) and 'n_samples' ({n_samples}) must be both greater than 0")
        sequential = self._concat_along_axis == 2
        seq_len = kwargs.pop("sequence_length", 1) if sequential else None
        # same draw as the reference (buffers.py:682): how many batch elements each environment contributes
        bs_per_buf = np.bincount(self._rng.integers(0, self._n_envs, (batch_size,)))
        T = seq_len if sequential else 1
        parts, next_parts = [], []
        for i, (b, bs) in enumerate(zip(self._buf, bs_per_buf)):
            if bs <= 0:
                continue
            if sequential:
                rows, nxt = b._plan(int(bs), sample_next_obs, n_samples, seq_len)
            else:
                rows, nxt = b._plan(int(bs), sample_next_obs, n_samples)
            if b.empty:
                raise RuntimeError("The buffer has not been initialized. Try to add some data first.")
            off = i * self._buffer_size
            parts.append((rows + off).reshape(n_samples, int(bs), T))
            if nxt is not None:
                next_parts.append((nxt + off).reshape(n_samples, int(bs), T))
        # concatenate the per-env plans along the batch axis -> one (sample, batch, time) index array
        rows = np.ascontiguousarray(np.concatenate(parts, axis=1)).reshape(-1)
        rows_d = torch.from_numpy(rows).to(self._device, non_blocking=True)
        next_d = None
        if next_parts:
            next_d = torch.from_numpy(np.ascontiguousarray(np.concatenate(next_parts, axis=1)).reshape(-1)).to(self._device)
        out = {}
        obs_keys = self._buf[0]._obs_keys
        for k in self._store:
            store = self._flat(k)
            shape = (n_samples, seq_len, batch_size) if sequential else (n_samples, batch_size)
            o = torch.empty((*shape, *store.shape[1:]), dtype=store.dtype, device=store.device)
            self.ops.replay_gather(store, rows_d, o, n_samples, batch_size, T)
            out[k] = o
            if sample_next_obs and (sequential or k in obs_keys):
                o2 = torch.empty_like(o)
                self.ops.replay_gather(store, next_d, o2, n_samples, batch_size, T)
                out[f"next_{k}"] = o2
        return _finish(out, dtype, device)

    def sample(self, batch_size: int, sample_next_obs: bool = False, clone: bool = False, n_samples: int = 1,
               **kwargs) -> Dict[str, np.ndarray]:
        t = self.sample_tensors(batch_size, sample_next_obs=sample_next_obs, n_samples=n_samples, **kwargs)
        return {k: v.cpu().numpy() for k, v in t.items()}
