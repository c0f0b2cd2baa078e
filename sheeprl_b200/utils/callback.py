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
    def __init__(self, keep_last: Optional[int] = None) -> None:
        self.keep_last = keep_last

    # ------------------------------------------------------------------ reference entry points
    def on_checkpoint_coupled(self, fabric, ckpt_path: str, state: Dict[str, Any], replay_buffer=None):
        if replay_buffer is not None:
            rb_state = self._ckpt_rb(replay_buffer)
            state["rb"] = replay_buffer
            if fabric.world_size > 1:
                # every rank's buffer goes into rank 0's file; host objects travel over gloo (callback.py:42-54)
                import torch.distributed as dist

                group = dist.new_group(backend="gloo")
                gathered = [None for _ in range(fabric.world_size)] if fabric.global_rank == 0 else None
                dist.gather_object(replay_buffer, gathered, dst=0, group=group)
                if fabric.global_rank == 0:
                    state["rb"] = gathered
                dist.destroy_process_group(group)
        fabric.save(ckpt_path, state)
        if replay_buffer is not None:
            self._experiment_consistent_rb(replay_buffer, rb_state)
        if fabric.is_global_zero and self.keep_last:
            self._delete_old_checkpoints(pathlib.Path(ckpt_path).parent)

    def on_checkpoint_player(self, fabric, player_trainer_collective, ckpt_path: str, replay_buffer=None,
                             ratio_state_dict: Optional[Dict[str, Any]] = None):
        state = [None]
        player_trainer_collective.broadcast_object_list(state, src=1)
        state = state[0]
        if replay_buffer is not None:
            rb_state = self._ckpt_rb(replay_buffer)
            state["rb"] = replay_buffer
        if ratio_state_dict is not None:
            state["ratio"] = ratio_state_dict
        fabric.save(ckpt_path, state)
        if replay_buffer is not None:
            self._experiment_consistent_rb(replay_buffer, rb_state)
        if fabric.is_global_zero and self.keep_last:
            self._delete_old_checkpoints(pathlib.Path(ckpt_path).parent)

    def on_checkpoint_trainer(self, fabric, player_trainer_collective, state: Dict[str, Any], ckpt_path: str):
        if fabric.global_rank == 1:
            player_trainer_collective.broadcast_object_list([state], src=1)
        fabric.save(ckpt_path, state)

    # ------------------------------------------------------------------ truncated fix-up
    @staticmethod
    def _rings(rb):
        if isinstance(rb, EnvIndependentReplayBuffer):
            return list(rb.buffer)
        if isinstance(rb, ReplayBuffer):
            return [rb]
        raise TypeError(f"unsupported replay buffer type {type(rb)} (EpisodeBuffer is not part of the B200 data path)")

    def _ckpt_rb(self, rb):
        saved = []
        for b in self._rings(rb):
            if b.empty or "truncated" not in b.buffer:
                saved.append(None)
                continue
            row = (b._pos - 1) % b.buffer_size
            saved.append(b["truncated"][row].clone())
            b["truncated"][row] = 1
        return saved

    def _experiment_consistent_rb(self, rb, saved) -> None:
        for b, old in zip(self._rings(rb), saved):
            if old is not None:
                b["truncated"][(b._pos - 1) % b.buffer_size] = old

    def _delete_old_checkpoints(self, ckpt_folder: pathlib.Path) -> None:
        ckpts = sorted(ckpt_folder.glob("*.ckpt"), key=os.path.getmtime)
        if len(ckpts) > self.keep_last:
            for f in ckpts[:-self.keep_last]:
                f.unlink()


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
