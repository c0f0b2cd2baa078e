"""Decoupled SAC data plane on CPU (BASELINE config 3 topology): world_size-3 gloo — rank 0 samples from a
replay ring and sends row blocks, ranks 1-2 train data-parallel on them, rank 1 returns the actor.  Checks that every
trainer got exactly its rows, that the trainers stay bit-identical, that the player ends up with their actor, and that
the result equals a single-process run of the same two-trainer average."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, W = 2, 3                                    # gradient steps per trainer and message, world size


def _noise(fx, rank, n):
    sp = fx["spec"]
    g = torch.Generator().manual_seed(1000 + rank)
    return [{"eps_next": torch.randn(sp["B"], sp["act_dim"], generator=g), "eps_cur": torch.randn(sp["B"], sp["act_dim"], generator=g)}
            for _ in range(n)]


def _rollout(fx, n):
    sp = fx["spec"]
    g = torch.Generator().manual_seed(7)
    return {"observations": torch.randn(n, 1, sp["obs_dim"], generator=g).numpy(),
            "next_observations": torch.randn(n, 1, sp["obs_dim"], generator=g).numpy(),
            "actions": torch.rand(n, 1, sp["act_dim"], generator=g).mul(2).sub(1).numpy(),
            "rewards": torch.randn(n, 1, 1, generator=g).numpy(),
            "terminated": (torch.rand(n, 1, 1, generator=g) < 0.1).float().numpy()}


def _player_sample(fx):
    import numpy as np

    from oracle.ops_emul import EmulOps
    from sheeprl_b200.data.buffers import ReplayBuffer

    rb = ReplayBuffer(64, 1, obs_keys=("observations",), device="cpu", ops=EmulOps())
    rb.add(_rollout(fx, 50))
    rb._rng = np.random.default_rng(3)
    s = rb.sample_tensors(G * fx["spec"]["B"] * (W - 1))
    return {k: v[0] for k, v in s.items()}          # [n_samples=1, rows, ...] -> [rows, ...]


def _worker(rank, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(W))
    torch.set_num_threads(2)
    from sheeprl_b200.algos.sac import decoupled as D
    from sheeprl_b200.parallel import attach_data_parallel, init_process_group_from_env
    from tests.test_sac_cpu import load, make_engine

    init_process_group_from_env("gloo")
    world, pair, optim = D.setup_groups()
    fx = load("sac_tiny")
    eng = make_engine(fx)
    if rank == 0:                                                  # ---- player (sac_decoupled.py:33-353)
        eng.actor.flat.zero_()                                     # stale weights: must be replaced by the trainers'
        D.broadcast_actor(eng, pair)                               # initial weights (:117-126)
        first = eng.actor.flat.clone()
        sample = _player_sample(fx)
        D.player_send_batch(sample, world)
        D.broadcast_actor(eng, pair)
        D.player_send_stop(world)
        out[0] = {"sample": sample, "first": first, "actor": eng.actor.flat.clone()}
    else:                                                          # ---- trainers (:356-544)
        attach_data_parallel(eng, optim)
        if rank == 1:
            D.broadcast_actor(eng, pair)
        got, updates = [], 0
        while True:
            data = D.trainer_recv_batch("cpu", world)
            if data is None:
                break
            got.append(data)
            B, nz = fx["spec"]["B"], _noise(fx, rank, G)
            updates += D.trainer_update(eng, {k: v[:B] for k, v in data.items()}, B, updates, 2, nz[:1])
            first_grad = eng.qf.grad.clone()                       # all-reduced over the two trainers
            updates += D.trainer_update(eng, {k: v[B:] for k, v in data.items()}, B, updates, 2, nz[1:])
            if rank == 1:
                D.broadcast_actor(eng, pair)
        out[rank] = {"data": got, "first_grad": first_grad, "actor": eng.actor.flat.clone(), "qf": eng.qf.flat.clone(), "alpha": eng.alpha.flat.clone(),
                     "updates": updates}
    dist.barrier()
    dist.destroy_process_group()


def test_player_trainers_round_trip():
    mp.set_start_method("spawn", force=True)
    out = mp.Manager().dict()
    mp.spawn(_worker, args=(31200 + (os.getpid() % 500), out), nprocs=W, join=True)
    p, t1, t2 = out[0], out[1], out[2]
    sys.path.insert(0, ROOT)
    from tests.test_sac_cpu import load, make_engine

    fx = load("sac_tiny")
    B = fx["spec"]["B"]
    rows = G * B
    for r, t in ((1, t1), (2, t2)):                                # each trainer received exactly its block, as float32
        assert len(t["data"]) == 1 and t["updates"] == G
        for k, v in p["sample"].items():
            assert torch.equal(t["data"][0][k], v[(r - 1) * rows: r * rows].float()), (r, k)
    for k in ("actor", "qf", "alpha"):
        assert torch.equal(t1[k], t2[k]), k                        # data-parallel replicas stay identical
    assert torch.equal(p["first"], make_engine(fx).actor.flat)     # the player started from the trainers' weights ...
    assert torch.equal(p["actor"], t1["actor"])                    # ... and ends with their updated actor
    assert not torch.equal(p["actor"], p["first"])
    # the critic gradient of the first update is the mean of the two trainers' single-process gradients
    grads = []
    for r in (1, 2):
        e = make_engine(fx)
        blk = {k: v[(r - 1) * rows: (r - 1) * rows + B].float() for k, v in p["sample"].items()}
        e.train_step(blk, True, _noise(fx, r, G)[0])
        grads.append(e.qf.grad.clone())
    want = 0.5 * (grads[0] + grads[1])
    assert float((t1["first_grad"] - want).abs().max()) <= 1e-5 * float(want.abs().max())
    assert torch.isfinite(t1["qf"]).all() and not torch.equal(make_engine(fx).qf.flat, t1["qf"])
