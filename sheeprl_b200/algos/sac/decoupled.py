"""Data plane of decoupled SAC (BASELINE config 3): rank 0 plays and owns the replay buffer, ranks 1..W-1
are data-parallel trainers.

Reference (`sheeprl/algos/sac/sac_decoupled.py`): the player pickles the sampled rows and `scatter_object_list`s one
chunk per trainer (:246-257), the trainers train on it minibatch by minibatch under DDP (:437-489), rank 1 broadcasts the
flattened actor back to the player (:259-263, :491-494); `-1` instead of a chunk ends the run (:343-351, :438-462).

Here the same three steps move TENSORS: a small pickled header (rows per trainer, key -> (shape, dtype); -1 rows = stop)
followed by point-to-point sends of contiguous row blocks straight from the device ring's gather output (NCCL
send/recv over NVLink on GPUs, gloo on CPU), and ONE broadcast of the actor's flat parameter group — the engine keeps it
flat, so there is no parameters_to_vector / vector_to_parameters round trip.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

STOP = -1


def setup_groups(world_size: Optional[int] = None):
    """(world group, player<->rank-1 pair, trainers' optimisation group) as in sac_decoupled.py:563-588; every rank must
    call this (new_group is collective)."""
    world = dist.get_world_size() if world_size is None else world_size
    if world < 2:
        raise RuntimeError("decoupled SAC needs at least two ranks: one player and one trainer")   # cli.py:291-308
    pair = dist.new_group(ranks=[0, 1])
    optim = dist.new_group(ranks=list(range(1, world)))
    return dist.group.WORLD, pair, optim


def _header(obj, group):
    box = [obj]
    dist.broadcast_object_list(box, src=0, group=group)
    return box[0]


def player_send_batch(sample: Dict[str, torch.Tensor], group=None, chunk_sizes: Optional[Sequence[int]] = None) -> None:
    """sample: {key: [n, ...]} (what `rb.sample_tensors` returned for all trainers, sac_decoupled.py:241-247).
    Trainer r receives the r-th block of rows of every key, as float32 (the reference sends `v.float()`): equal blocks
    by default, `chunk_sizes` (one per trainer) for PPO's near-even split (ppo_decoupled.py:293-299)."""
    world = dist.get_world_size(group)
    n = next(iter(sample.values())).shape[0]
    if chunk_sizes is None:
        if n % (world - 1):
            raise ValueError(f"{n} sampled rows do not split evenly over {world - 1} trainers")
        chunk_sizes = [n // (world - 1)] * (world - 1)
    if len(chunk_sizes) != world - 1 or sum(chunk_sizes) != n:
        raise ValueError(f"chunk_sizes {list(chunk_sizes)} do not cover {n} rows for {world - 1} trainers")
    data = {k: (v if v.dtype == torch.float32 else v.float()).contiguous() for k, v in sample.items()}
    _header({"rows": list(chunk_sizes), "spec": {k: tuple(v.shape[1:]) for k, v in data.items()}}, group)
    starts = [sum(chunk_sizes[:i]) for i in range(world - 1)]
    ops = [dist.P2POp(dist.isend, v[starts[r - 1]: starts[r - 1] + chunk_sizes[r - 1]], r, group)
           for r in range(1, world) for v in data.values()]
    for req in dist.batch_isend_irecv(ops):
        req.wait()


def player_send_stop(group=None) -> None:
    _header({"rows": STOP, "spec": {}}, group)


def trainer_recv_batch(device, group=None) -> Optional[Dict[str, torch.Tensor]]:
    """blocks until the player's next message; None when the player sent the stop marker"""
    head = _header(None, group)
    if head["rows"] == STOP:
        return None
    rows = head["rows"][dist.get_rank(group) - 1]
    out = {k: torch.empty((rows, *tail), dtype=torch.float32, device=device) for k, tail in head["spec"].items()}
    for req in dist.batch_isend_irecv([dist.P2POp(dist.irecv, v, 0, group) for v in out.values()]):
        req.wait()
    return out


def broadcast_actor(engine, pair_group) -> None:
    """rank 1 -> rank 0: the actor's flat parameter group (both ends call it; other trainers must not)"""
    dist.broadcast(engine.actor.flat, src=1, group=pair_group)


def broadcast_flat(flat: torch.Tensor, pair_group) -> None:
    """rank 1 -> rank 0: any flat parameter group — decoupled PPO returns the WHOLE agent (ppo_decoupled.py:302-305,
    :551-556), which `PPOEngine` keeps as one flat group"""
    dist.broadcast(flat, src=1, group=pair_group)


def minibatches(n_rows: int, batch_size: int) -> Sequence[Tuple[int, int]]:
    """the trainers' BatchSampler(range(n), batch_size, drop_last=False) (sac_decoupled.py:464-466) as row ranges"""
    return [(s, min(s + batch_size, n_rows)) for s in range(0, n_rows, batch_size)]


def trainer_update(engine, data: Dict[str, torch.Tensor], batch_size: int, first_update: int, ema_every: int,
                   noise: Optional[Sequence[Dict[str, torch.Tensor]]] = None) -> int:
    """the trainer's inner loop (sac_decoupled.py:472-489): one engine update per minibatch of the received chunk; the
    gradient all-reduces over the trainers are the hooks `attach_data_parallel(engine, optimisation_group)` installed.
    Returns the number of updates done."""
    n = next(iter(data.values())).shape[0]
    for i, (s, e) in enumerate(minibatches(n, batch_size)):
        if e - s != engine.B:
            engine.B = e - s
            engine._alloc()
        engine.train_step({k: v[s:e] for k, v in data.items()}, (first_update + i) % ema_every == 0,
                          None if noise is None else noise[i])
    return len(minibatches(n, batch_size))
