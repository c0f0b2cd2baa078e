"""Data plane of decoupled PPO (reference `sheeprl/algos/ppo/ppo_decoupled.py`): rank 0 collects the rollout, permutes
and splits it near-evenly over the trainers (:293-299), ranks 1..W-1 run `update_epochs` passes of minibatches under data
parallelism (:486-548) and rank 1 returns the WHOLE agent as one flat vector (:302-305, :551-556).

Transfers are the tensor collectives of `sheeprl_b200.algos.sac.decoupled` (header + point-to-point row blocks, one
broadcast of the engine's flat group).  Trainers may hold different numbers of minibatches; the reference wraps its loop in
DDP's `Join` (:497-499): a trainer that ran out keeps answering the others' all-reduces with zero gradients (the average
still divides by the number of trainers) and at the end every trainer takes the parameters of the highest-ranked trainer
among those that ran longest.  `trainer_update` reproduces exactly that on the flat group.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.distributed as dist
from torch.utils.data import BatchSampler, RandomSampler

from sheeprl_b200.algos.sac.decoupled import (broadcast_flat, player_send_batch, player_send_stop, setup_groups,  # noqa: F401
                                              trainer_recv_batch)


def chunk_sizes(n_rows: int, n_trainers: int) -> List[int]:
    """sizes of `torch.tensor_split(arange(n_rows), n_trainers)`: the first n_rows % n_trainers chunks get one more row"""
    base, extra = divmod(n_rows, n_trainers)
    return [base + (1 if i < extra else 0) for i in range(n_trainers)]


def player_send_rollout(local_data: Dict[str, torch.Tensor], group=None, generator: Optional[torch.Generator] = None) -> None:
    """local_data: flat `[N, ...]` rollout tensors (returns / advantages included).  A random permutation of the rows is
    split over the trainers (ppo_decoupled.py:293-299)."""
    n = next(iter(local_data.values())).shape[0]
    dev = next(iter(local_data.values())).device
    perm = torch.randperm(n, generator=generator).to(dev)
    player_send_batch({k: v[perm] for k, v in local_data.items()}, group, chunk_sizes(n, dist.get_world_size(group) - 1))


def index_batches(n_rows: int, batch_size: int, epochs: int) -> List[List[int]]:
    """the trainer's sampler stack: BatchSampler(RandomSampler(range(n)), batch_size, drop_last=False), re-drawn every
    epoch (ppo_decoupled.py:486-500)"""
    sampler = BatchSampler(RandomSampler(range(n_rows)), batch_size=batch_size, drop_last=False)
    return [list(b) for _ in range(epochs) for b in sampler]


def trainer_update(engine, data: Dict[str, torch.Tensor], batches: Sequence[Sequence[int]], optim_group, on_minibatch=None) -> int:
    """Runs this trainer's minibatches on `engine` (whose all-reduce hook works on `optim_group`) with DDP-Join semantics
    for uneven counts.  Returns the number of optimisation steps the group took."""
    # bookkeeping tensors live on the engine's device: `optim_group` inherits the default backend (NCCL on GPUs), which
    # has no CPU transport
    mine = torch.tensor([len(batches)], dtype=torch.int64, device=engine.device)
    counts = [torch.zeros(1, dtype=torch.int64, device=engine.device) for _ in range(dist.get_world_size(optim_group))]
    dist.all_gather(counts, mine, group=optim_group)
    counts = [int(c) for c in counts]
    longest = max(counts)
    engine.train(data, batches, on_minibatch)
    if engine.allreduce is not None:
        zeros = torch.zeros_like(engine.group.grad)
        for _ in range(longest - len(batches)):            # joined: shadow the others' collectives with zero gradients
            zeros.zero_()
            engine.allreduce(zeros, "agent")
    if min(counts) != longest:
        # Join's final model sync: parameters of the highest-ranked trainer among the last to finish
        src_local = max(i for i, c in enumerate(counts) if c == longest)
        dist.broadcast(engine.group.flat, src=dist.get_global_rank(optim_group, src_local), group=optim_group)
    return longest
