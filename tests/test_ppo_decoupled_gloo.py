"""Decoupled PPO data plane on CPU: world_size-3 gloo — rank 0 permutes / splits a 33-row rollout 17 + 16 over two
trainers (minibatch 16 => 2 vs 1 minibatches: the uneven case DDP's Join handles in the reference), the trainers train
data-parallel, rank 1 returns the whole agent.  Checks rows, Join semantics and the final weights on the player."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W, N_ROWS, BATCH = 3, 33, 16


def _worker(rank, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(W))
    torch.set_num_threads(2)
    from sheeprl_b200.algos.ppo import decoupled as D
    from sheeprl_b200.parallel import attach_data_parallel, init_process_group_from_env
    from tests.test_ppo_cpu import load, make_engine

    init_process_group_from_env("gloo")
    world, pair, optim = D.setup_groups()
    fx = load("ppo_vector")
    eng = make_engine(fx)
    if rank == 0:
        eng.group.flat.zero_()
        rollout = {k: v[:N_ROWS].clone() for k, v in fx["data"].items()}
        D.player_send_rollout(rollout, world, torch.Generator().manual_seed(5))
        D.broadcast_flat(eng.group.flat, pair)
        D.player_send_stop(world)
        out[0] = {"rollout": rollout, "agent": eng.group.flat.clone()}
    else:
        attach_data_parallel(eng, optim)
        data = D.trainer_recv_batch("cpu", world)
        n = data["actions"].shape[0]
        batches = [list(range(s, min(s + BATCH, n))) for s in range(0, n, BATCH)]       # deterministic for the check
        grads = []
        steps = D.trainer_update(eng, data, batches, optim, on_minibatch=lambda l: grads.append(eng.group.grad.clone()))
        if rank == 1:
            D.broadcast_flat(eng.group.flat, pair)
        assert D.trainer_recv_batch("cpu", world) is None
        out[rank] = {"data": data, "agent": eng.group.flat.clone(), "steps": steps, "grads": grads, "n_batches": len(batches)}
    dist.barrier()
    dist.destroy_process_group()


def test_uneven_trainers_follow_join_semantics():
    mp.set_start_method("spawn", force=True)
    out = mp.Manager().dict()
    mp.spawn(_worker, args=(31700 + (os.getpid() % 500), out), nprocs=W, join=True)
    p, t1, t2 = out[0], out[1], out[2]
    sys.path.insert(0, ROOT)
    from sheeprl_b200.algos.ppo.decoupled import chunk_sizes
    from tests.test_ppo_cpu import load, make_engine

    assert chunk_sizes(N_ROWS, 2) == [17, 16] and (t1["n_batches"], t2["n_batches"], t1["steps"]) == (2, 1, 2)
    perm = torch.randperm(N_ROWS, generator=torch.Generator().manual_seed(5))
    for k, v in p["rollout"].items():                               # permuted rows, split 17 + 16, float32
        assert torch.equal(t1["data"][k], v[perm][:17].float()) and torch.equal(t2["data"][k], v[perm][17:].float()), k
    assert torch.equal(t1["agent"], t2["agent"]) and torch.equal(p["agent"], t1["agent"])
    # step 1: mean of the two trainers' gradients; step 2: trainer 2 has joined -> half of trainer 1's gradient
    fx = load("ppo_vector")
    e1, e2 = make_engine(fx), make_engine(fx)
    g = {}
    e1.allreduce = lambda flat, name: g.__setitem__("a", flat.clone())
    e2.allreduce = lambda flat, name: g.__setitem__("b", flat.clone())
    e1.minibatch_step(t1["data"], torch.arange(0, 16))
    e2.minibatch_step(t2["data"], torch.arange(0, 16))
    want1 = 0.5 * (g["a"] + g["b"])
    assert float((t1["grads"][0] - want1).abs().max()) <= 1e-5 * float(want1.abs().max())
    assert float((t2["grads"][0] - want1).abs().max()) <= 1e-5 * float(want1.abs().max())
    assert float(t1["grads"][1].abs().max()) > 0                    # the second step ran on trainer 1 only ...
    assert len(t2["grads"]) == 1                                    # ... trainer 2 shadowed it and took trainer 1's weights
