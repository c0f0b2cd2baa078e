"""Checkpoint callback for the device-resident buffers — the reference's `CheckpointCallback`
(sheeprl/utils/callback.py:14-150) on `sheeprl_b200.data.buffers`: same three entry points, same on-disk content
(`state["rb"]` is the pickled buffer object, a list of them when world_size > 1), same keep_last pruning.

The environments' state is not part of a checkpoint, so the last written step of every ring is marked `truncated`
while the file is written and restored afterwards (callback.py:87-142)."""
from __future__ import annotations

import os
import pathlib
from typing import Any, Dict, Optional

from sheeprl_b200.data.buffers import EnvIndependentReplayBuffer, ReplayBuffer


class CheckpointCallback:
    """Hooks `fabric.call("on_checkpoint_*", ...)` reaches (same names and keyword arguments as the reference's callback).
    All three funnel into `_write`: mark the newest step of every ring as truncated, put the buffer(s) into the state,
    save, undo the mark, prune old files."""

    def __init__(self, keep_last: Optional[int] = None) -> None:
        self.keep_last = keep_last

    # ------------------------------------------------------------------ one writer behind the three hooks
    def _write(self, fabric, path: str, state: Dict[str, Any], rb, gather: bool) -> None:
        rings = self._rings(rb) if rb is not None else []
        marks = self._mark_truncated(rings)
        try:
            if rb is not None:
                state["rb"] = self._all_ranks(fabric, rb) if gather else rb
            fabric.save(path, state)
        finally:
            for ring, old in zip(rings, marks):               # the run goes on with the episode it was in
                if old is not None:
                    ring["truncated"][(ring._pos - 1) % ring.buffer_size] = old
        if self.keep_last and fabric.is_global_zero:
            files = sorted(pathlib.Path(path).parent.glob("*.ckpt"), key=os.path.getmtime)
            for stale in files[: max(0, len(files) - self.keep_last)]:
                stale.unlink()

    @staticmethod
    def _all_ranks(fabric, rb):
        """world_size > 1: rank 0's file holds the list of every rank's buffer (host objects travel over a gloo group)"""
        import torch.distributed as dist

        group = dist.new_group(backend="gloo")
        box = [None] * fabric.world_size if fabric.global_rank == 0 else None
        dist.gather_object(rb, box, dst=0, group=group)
        dist.destroy_process_group(group)
        return box if fabric.global_rank == 0 else rb

    def on_checkpoint_coupled(self, fabric, ckpt_path: str, state: Dict[str, Any], replay_buffer=None):
        self._write(fabric, ckpt_path, state, replay_buffer, gather=fabric.world_size > 1)

    def on_checkpoint_player(self, fabric, player_trainer_collective, ckpt_path: str, replay_buffer=None,
                             ratio_state_dict: Optional[Dict[str, Any]] = None):
        box = [None]
        player_trainer_collective.broadcast_object_list(box, src=1)       # the model state lives on the first trainer
        state = box[0]
        if ratio_state_dict is not None:
            state["ratio"] = ratio_state_dict
        self._write(fabric, ckpt_path, state, replay_buffer, gather=False)

    def on_checkpoint_trainer(self, fabric, player_trainer_collective, state: Dict[str, Any], ckpt_path: str):
        if fabric.global_rank == 1:
            player_trainer_collective.broadcast_object_list([state], src=1)
        fabric.save(ckpt_path, state)

    # ------------------------------------------------------------------ truncated mark
    @staticmethod
    def _rings(rb):
        if isinstance(rb, EnvIndependentReplayBuffer):
            return list(rb.buffer)
        if isinstance(rb, ReplayBuffer):
            return [rb]
        raise TypeError(f"unsupported replay buffer type {type(rb)} (EpisodeBuffer is not part of the B200 data path)")

    @staticmethod
    def _mark_truncated(rings):
        """The environments' state is not checkpointed: a resumed run starts new episodes, so the newest stored step of
        every ring must read as the end of one.  Returns the overwritten values."""
        old = []
        for ring in rings:
            if ring.empty or "truncated" not in ring.buffer:
                old.append(None)
                continue
            row = (ring._pos - 1) % ring.buffer_size
            old.append(ring["truncated"][row].clone())
            ring["truncated"][row] = 1
        return old


def load_replay_buffer(obj, device="cuda", ops=None):
    """`state["rb"]` of a checkpoint -> device buffer(s): objects of this package pass through (re-bound to `ops`),
    objects of the reference's classes (checkpoints written by sheeprl itself) are adopted via `from_reference`."""
    import sheeprl_b200.data.buffers as B

    if isinstance(obj, (list, tuple)):
        return [load_replay_buffer(o, device, ops) for o in obj]
    if isinstance(obj, (B.ReplayBuffer, B.EnvIndependentReplayBuffer)):
        if ops is not None:
            obj._ops = ops
            for b in (obj.buffer if isinstance(obj, B.EnvIndependentReplayBuffer) else ()):
                b._ops = ops
        return obj
    name = type(obj).__name__
    if name not in ("ReplayBuffer", "SequentialReplayBuffer", "EnvIndependentReplayBuffer"):
        raise TypeError(f"cannot adopt a replay buffer of type {type(obj)}")
    return getattr(B, name).from_reference(obj, device=device, ops=ops)
