"""CUDA-graph replay of an update step.

An engine's `train_step` is a fixed schedule of kernel launches (and, under data parallelism, NCCL collectives) on
device-resident buffers: nothing on it allocates, synchronises or branches on device data (the Adam step count and the
Philox counters live on the device).  `StepGraph` captures that schedule once per input signature and replays it on every
later call, which removes the ~450 host-side launches of one Dreamer-V3 update from the critical path
(`train()` in `sheeprl_b200/algos/dreamer_v3/dreamer_v3.py` goes through here by default).

The caller's batch is copied into static input buffers first (device-to-device, or host-to-device when the batch still
lives in pinned host memory), so the reference's calling convention — a fresh dict of tensors per call — is unchanged.
"""
from __future__ import annotations

import atexit
import weakref
from typing import Callable, Dict, Optional, Tuple

import torch

_LIVE = weakref.WeakSet()


def release_all() -> None:
    """Drops every captured graph of the process.  A graph that contains NCCL kernels keeps the communicator busy:
    `torch.distributed.destroy_process_group()` (and NCCL's teardown at interpreter exit) blocks until the graph is gone,
    so graphs must die first — at exit (registered below, runs before torch's own handlers) or explicitly before
    destroying the process group."""
    for g in list(_LIVE):
        g.reset()
    if torch.cuda.is_available():
        try:
            torch.cuda.synchronize()
        except Exception:
            pass


atexit.register(release_all)


class _Entry:
    __slots__ = ("static", "graph", "calls")

    def __init__(self, static):
        self.static, self.graph, self.calls = static, None, 0


class StepGraph:
    def __init__(self, device, warmup: int = 2, on_replay: Optional[Callable[[], None]] = None,
                 host_state: Optional[Tuple[Callable[[], object], Callable[[object], None]]] = None):
        """warmup: eager calls before the capture (lazy allocations, cuBLAS/NCCL initialisation).
        on_replay: called after every replay (host mirrors of device-side counters).
        host_state: (save, restore) around the capture pass — capturing runs the host code of the step without
        executing its kernels, so host mirrors touched by it must be put back."""
        self.device = torch.device(device)
        self.warmup = int(warmup)
        self.on_replay = on_replay
        self.host_state = host_state
        self._entries: Dict[tuple, _Entry] = {}
        _LIVE.add(self)

    @staticmethod
    def signature(data: Dict[str, torch.Tensor]) -> tuple:
        return tuple((k, tuple(v.shape), v.dtype) for k, v in sorted(data.items()))

    def reset(self) -> None:
        """Drops every captured graph (call after changing anything the capture baked in: hyper-parameters passed by
        value such as the learning rate, the data-parallel hooks, the kernel choice)."""
        self._entries.clear()

    def captured(self, data: Dict[str, torch.Tensor], key: tuple = ()) -> bool:
        e = self._entries.get(self.signature(data) + tuple(key))
        return e is not None and e.graph is not None

    def run(self, step: Callable[[Dict[str, torch.Tensor]], None], data: Dict[str, torch.Tensor], key: tuple = ()) -> None:
        """key: values the capture bakes in besides the input shapes (e.g. the learning rates): a change re-captures."""
        sig = self.signature(data) + tuple(key)
        e = self._entries.get(sig)
        if e is None:
            e = _Entry({k: torch.empty(v.shape, dtype=v.dtype, device=self.device) for k, v in data.items()})
            self._entries[sig] = e
        for k, v in data.items():
            e.static[k].copy_(v, non_blocking=True)
        if e.graph is not None:
            e.graph.replay()
            if self.on_replay is not None:
                self.on_replay()
            return
        if e.calls < self.warmup:
            e.calls += 1
            step(e.static)
            return
        saved = self.host_state[0]() if self.host_state is not None else None
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            step(e.static)
        if self.host_state is not None:
            self.host_state[1](saved)
        e.graph = g
        g.replay()
        if self.on_replay is not None:
            self.on_replay()
