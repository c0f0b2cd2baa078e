"""Hardware runs of the two multi-GPU configurations of BASELINE.json that bench.py does not cover:

    C3  PPO pixel 84x84 obs, 16 vec-envs per rank, N x B200 data-parallel        (ppo_dp)
    C4  SAC-decoupled continuous 17-dim obs, 1M replay buffer, 1 player + (N-1) trainers   (sac_decoupled)

Launch (one process per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        tools/run_multi_gpu.py ppo_dp --calls 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 \
        tools/run_multi_gpu.py sac_decoupled --messages 40

Both drive the product's own data planes (`parallel.attach_data_parallel`, `algos/sac/decoupled.py`) on synthetic data of
the configuration's shape and print ONE JSON line on rank 0: device-timed (CUDA events, max over ranks) throughput.
References: ppo.py:30-102 / :373 (train per iteration), sac_decoupled.py:241-263, 437-494, 563-588.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class _Event:
    """CUDA event on a GPU, wall clock in the --cpu-emulation dry run"""

    def __init__(self, dev):
        import time

        self.cuda = torch.device(dev).type == "cuda"
        self.ev = torch.cuda.Event(enable_timing=True) if self.cuda else None
        self.t, self._time = 0.0, time

    def record(self):
        if self.cuda:
            self.ev.record()
        else:
            self.t = self._time.perf_counter()

    def elapsed_time(self, other):
        return self.ev.elapsed_time(other.ev) if self.cuda else (other.t - self.t) * 1e3


def _sync(dev):
    if torch.device(dev).type == "cuda":
        torch.cuda.synchronize()


def _ops(args, dev):
    if args.cpu_emulation:
        from oracle.ops_emul import EmulOps      # dry run of the script's control flow on a GPU-less host

        return EmulOps()
    from sheeprl_b200.lib import CudaOps

    return CudaOps(dev)


def _max_over_ranks(ms: float, dev, group=None) -> float:
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t[0])


def ppo_dp(args, rank, world, dev):
    """C3: every rank holds its own rollout (16 envs x 128 steps = 2048 rows of 12x84x84 uint8) and runs the reference's
    update loop (10 epochs x 32 minibatches of 64) through the public train(); gradients all-reduced per minibatch."""
    from sheeprl_b200.algos.ppo.agent import default_init
    from sheeprl_b200.algos.ppo.engine import PPOEngine
    from sheeprl_b200.parallel import attach_data_parallel

    cu = _ops(args, dev)
    spec = dict(cnn_channels=12, screen=84, mlp_dim=0, dense=64, layers=2, cnn_features=512, mlp_features=64,
                actions_dim=(6,), is_continuous=False, act="tanh")
    hp = dict(clip_coef=0.2, vf_coef=1.0, ent_coef=0.01, clip_vloss=False, normalize_advantages=True, max_grad_norm=0.5)
    opt = {"lr": 1e-3, "eps": 1e-4, "betas": (0.9, 0.999)}
    eng = PPOEngine(spec, hp, opt, dev, cu)
    eng.load_reference_state(default_init(eng.reference_shapes(), torch.Generator().manual_seed(0)))
    attach_data_parallel(eng)
    n_envs, rollout, mb, epochs = (16, 128, 64, 10) if not args.cpu_emulation else (2, 8, 4, 1)
    N = n_envs * rollout
    g = torch.Generator().manual_seed(100 + rank)
    data = {"rgb": torch.randint(0, 256, (N, 12, 84, 84), generator=g, dtype=torch.uint8).to(dev),
            "actions": torch.nn.functional.one_hot(torch.randint(0, 6, (N,), generator=g), 6).float().to(dev),
            "logprobs": (-torch.rand(N, 1, generator=g) * 2).to(dev), "values": torch.randn(N, 1, generator=g).to(dev),
            "advantages": torch.randn(N, 1, generator=g).to(dev)}
    data["returns"] = data["values"] + 0.5 * torch.randn(N, 1, generator=g).to(dev)

    def batches():
        out = []
        for _ in range(epochs):
            perm = torch.randperm(N, generator=g)
            out += [perm[i: i + mb].tolist() for i in range(0, N, mb)]
        return out

    l0 = getattr(cu, "launches", 0)
    eng.train(data, batches())                                   # warm-up call (also counts the launches)
    launches = getattr(cu, "launches", 0) - l0
    _sync(dev)
    dist.barrier()
    e0, e1 = _Event(dev), _Event(dev)
    e0.record()
    for _ in range(args.calls):
        eng.train(data, batches())
    e1.record()
    _sync(dev)
    ms = _max_over_ranks(e0.elapsed_time(e1), dev)
    flat = eng.group.flat.clone()
    ref = flat.clone()
    dist.broadcast(ref, src=0)
    same = bool(torch.equal(flat, ref))
    ok = torch.tensor([int(same)], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        n_mb = epochs * (N // mb)
        return ({
            "config": "C3: PPO pixel 12x84x84 uint8, 16 envs x 128 steps per rank, minibatch 64, 10 epochs, NatureCNN-512",
            "n_gpus": world, "metric": "PPO minibatch updates/s (whole job)", "value": world * args.calls * n_mb / (ms / 1e3),
            "unit": "minibatches/s", "train_calls_per_s": args.calls / (ms / 1e3), "ms_per_train_call": ms / args.calls,
            "ms_per_minibatch": ms / args.calls / n_mb, "env_steps_per_s": world * args.calls * N / (ms / 1e3),
            "allreduce_bytes_per_minibatch": int(eng.group.numel * 4), "gpu_launches_per_train_call": launches,
            "replicas_identical": bool(int(ok[0])), "losses_finite": bool(torch.isfinite(eng.losses).all()),
            "scaling": "weak"})
    return None


def sac_decoupled(args, rank, world, dev):
    """C4: rank 0 = player with a 1M-row device ring (17-dim obs, 6-dim actions), ranks 1..W-1 = data-parallel trainers.
    Per message the player gathers G x 256 rows per trainer from the ring and sends the row blocks point-to-point; the
    trainers run G updates each (gradients all-reduced over the trainers' group); rank 1 broadcasts the actor back."""
    import numpy as np

    from sheeprl_b200.algos.sac import decoupled as D
    from sheeprl_b200.algos.sac.engine import SACEngine
    from sheeprl_b200.data.buffers import ReplayBuffer
    from sheeprl_b200.parallel import attach_data_parallel

    cu = _ops(args, dev)
    O, A, H, B, G = (17, 6, 256, 256, args.gradient_steps) if not args.cpu_emulation else (5, 2, 16, 8, 2)
    opt = {"lr": 3e-4, "eps": 1e-4, "betas": (0.9, 0.999)}
    wg, pair, optim = D.setup_groups()
    eng = SACEngine(O, A, H, H, 2, B, 0.99, 0.005, 1.0, -1.0, 1.0, opt, opt, opt, dev, cu, seed=7)
    gen = torch.Generator().manual_seed(0)
    with torch.no_grad():
        for grp in (eng.actor, eng.qf):
            grp.flat.copy_((torch.rand(grp.numel, generator=gen) * 2 - 1).mul(0.05).to(dev))
        eng.qf_target.flat.copy_(eng.qf.flat)
    trainers = world - 1
    e0, e1 = _Event(dev), _Event(dev)
    if rank == 0:
        size = args.buffer_size
        rb = ReplayBuffer(size, 1, obs_keys=("observations", "next_observations"), device=dev, ops=cu)
        rng = np.random.default_rng(1)
        chunk = 100_000
        for s in range(0, size, chunk):
            n = min(chunk, size - s)
            rb.add({"observations": rng.standard_normal((n, 1, O), dtype=np.float32),
                    "next_observations": rng.standard_normal((n, 1, O), dtype=np.float32),
                    "actions": rng.uniform(-1, 1, (n, 1, A)).astype(np.float32),
                    "rewards": rng.standard_normal((n, 1, 1), dtype=np.float32),
                    "terminated": (rng.random((n, 1, 1)) < 0.01).astype(np.float32)})
        D.broadcast_actor(eng, pair)
        for it in range(args.warmup + args.messages):
            if it == args.warmup:
                _sync(dev)
                dist.barrier()
                e0.record()
            s = rb.sample_tensors(G * B * trainers)
            D.player_send_batch({k: v[0] for k, v in s.items()}, wg)
            D.broadcast_actor(eng, pair)
        e1.record()
        D.player_send_stop(wg)
        buf_bytes = sum(v.numel() * v.element_size() for v in rb._buf.values())
    else:
        attach_data_parallel(eng, optim)
        if rank == 1:
            D.broadcast_actor(eng, pair)
        updates, it = 0, 0
        while True:
            if it == args.warmup:
                _sync(dev)
                dist.barrier()
                e0.record()
            data = D.trainer_recv_batch(dev, wg)
            if data is None:
                break
            updates += D.trainer_update(eng, data, B, updates, 2)
            if rank == 1:
                D.broadcast_actor(eng, pair)
            it += 1
        e1.record()
    _sync(dev)
    ms = _max_over_ranks(e0.elapsed_time(e1), dev)
    # trainers must hold identical replicas; the player must hold the trainers' actor
    a = eng.actor.flat.clone()
    ref = a.clone()
    dist.broadcast(ref, src=1)
    ok = torch.tensor([int(torch.equal(a, ref))], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        per_msg = G * trainers
        return ({
            "config": f"C4: SAC decoupled, obs17 act6 hidden256 batch256, {args.buffer_size} row device ring on the player, "
                      f"{trainers} data-parallel trainers, {G} gradient steps per trainer and message",
            "n_gpus": world, "metric": "SAC gradient steps/s (sum over trainers)", "value": args.messages * per_msg / (ms / 1e3),
            "unit": "updates/s", "group_updates_per_s": args.messages * G / (ms / 1e3), "ms_per_message": ms / args.messages,
            "rows_sent_per_message": G * B * trainers, "replay_ring_bytes": int(buf_bytes),
            "actor_broadcast_bytes": int(eng.actor.numel * 4),
            "allreduce_bytes_per_update": int((eng.actor.numel + eng.qf.numel + eng.alpha.numel) * 4),
            "actor_identical_on_all_ranks": bool(int(ok[0])), "losses_finite": True})
    assert torch.isfinite(eng.metrics).all(), "non-finite SAC losses"
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["ppo_dp", "sac_decoupled"])
    ap.add_argument("--calls", type=int, default=3)
    ap.add_argument("--messages", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--gradient-steps", type=int, default=8)
    ap.add_argument("--buffer-size", type=int, default=1_000_000)
    ap.add_argument("--cpu-emulation", action="store_true", help="dry run of the control flow: gloo, CPU, torch test double")
    args = ap.parse_args()
    from sheeprl_b200.parallel import init_process_group_from_env

    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)                       # NCCL banners go to stderr; the JSON line is the only stdout output
    rank, local, world = init_process_group_from_env("gloo" if args.cpu_emulation else "nccl")
    assert world >= 2, "launch with torchrun on >= 2 GPUs"
    dev = torch.device("cpu") if args.cpu_emulation else torch.device("cuda", local)
    if not args.cpu_emulation:
        torch.cuda.set_device(dev)
    try:
        out = {"ppo_dp": ppo_dp, "sac_decoupled": sac_decoupled}[args.what](args, rank, world, dev)
        if out is not None:
            os.write(saved, (json.dumps(out) + "\n").encode())
    finally:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
