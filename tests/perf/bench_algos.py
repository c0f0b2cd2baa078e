"""(lives under tests/: it times the oracle as the CPU baseline, which only tests/, smoke() and bench.py may import)
Secondary measurements (not the headline bench.py line): SAC updates/s (BASELINE config 3 shapes: 17-dim obs, 6-dim
action, 256 hidden, batch 256) and PPO minibatches/s (config 2 shapes: 64 x 12x84x84 uint8, NatureCNN-512), each as a
CUDA graph of the engine's launches, next to the oracle (torch CPU) on the host cores.

    python tests/perf/bench_algos.py [--steps 200] [--no-cpu]        -> one JSON line per algorithm
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def timed_graph(step, steps, warmup=5):
    for _ in range(warmup):
        step()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        step()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        step()
    for _ in range(5):
        g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def timed_eager(step, steps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def bench_sac(args, cu):
    from oracle import sac_oracle as SO
    from oracle.dv3_oracle import AdamState
    from sheeprl_b200.algos.sac.engine import SACEngine

    O, A, H, B = 17, 6, 256, 256
    opt = {"lr": 3e-4, "eps": 1e-4, "betas": (0.9, 0.999)}
    eng = SACEngine(O, A, H, H, 2, B, 0.99, 0.005, 1.0, -1.0, 1.0, opt, opt, opt, "cuda", cu)
    P = SO.init_params(O, A, H, 2, seed=0)
    eng.load_reference_state(P["actor"], P["qf"], P["qf_target"], P["log_alpha"]["log_alpha"])
    data = {k: v.cuda() for k, v in SO.make_batch(B, O, A, seed=1).items()}
    l0 = cu.launches
    eng.train_step(data, True)
    launches = cu.launches - l0
    ms_graph = timed_graph(lambda: eng.train_step(data, True), args.steps)
    ms_eager = timed_eager(lambda: eng.train_step(data, True), args.steps)
    assert torch.isfinite(eng.metrics).all()
    cpu = None
    if args.cpu:
        torch.set_num_threads(min(16, os.cpu_count() or 1))
        opts = [AdamState(P[g], 3e-4, 1e-4) for g in ("qf", "actor", "log_alpha")]
        cd = SO.make_batch(B, O, A, seed=1)
        sc, bi = torch.ones(A), torch.zeros(A)
        f = lambda: SO.sac_train_step(P, *opts, cd, torch.randn(B, A), torch.randn(B, A), 0.99, 0.005, True, 2, sc, bi, -float(A))  # noqa
        for _ in range(5):
            f()
        t0 = time.perf_counter()
        n = 100
        for _ in range(n):
            f()
        cpu = {"value": n / (time.perf_counter() - t0), "unit": "updates/s", "cores": torch.get_num_threads(), "kind": "port"}
    print(json.dumps({"metric": "SAC updates/s (obs17 act6 hidden256 batch256, twin critics)", "value": 1e3 / ms_graph,
                      "unit": "updates/s", "us_per_update_graph": ms_graph * 1e3, "us_per_update_eager": ms_eager * 1e3,
                      "gpu_launches_per_update": launches, "cpu_baseline": cpu}))


def bench_ppo(args, cu):
    from oracle import ppo_oracle as PO
    from oracle.dv3_oracle import AdamState
    from sheeprl_b200.algos.ppo.agent import default_init
    from sheeprl_b200.algos.ppo.engine import PPOEngine

    spec = dict(cnn_channels=12, screen=84, mlp_dim=0, dense=64, layers=2, cnn_features=512, mlp_features=64,
                actions_dim=(6,), is_continuous=False, act="tanh")
    hp = dict(clip_coef=0.2, vf_coef=1.0, ent_coef=0.01, clip_vloss=False, normalize_advantages=False, max_grad_norm=0.0)
    opt = {"lr": 1e-3, "eps": 1e-4, "betas": (0.9, 0.999)}
    eng = PPOEngine(spec, hp, opt, "cuda", cu)
    init = default_init(eng.reference_shapes(), torch.Generator().manual_seed(0))
    eng.load_reference_state(init)
    N, B = 2048, 64                                       # rollout 128 steps x 16 envs, minibatch 64
    data = PO.make_rollout(spec, N, seed=1)
    dev = {k: v.cuda() for k, v in data.items()}
    dev["rgb"] = dev["rgb"].to(torch.uint8)
    idx = torch.randperm(N)[:B].cuda()
    l0 = cu.launches
    eng.minibatch_step(dev, idx)
    launches = cu.launches - l0
    ms_graph = timed_graph(lambda: eng.minibatch_step(dev, idx), args.steps)
    ms_eager = timed_eager(lambda: eng.minibatch_step(dev, idx), args.steps)
    assert torch.isfinite(eng.losses).all()
    cpu = None
    if args.cpu:
        torch.set_num_threads(min(16, os.cpu_count() or 1))
        p = {k: v.clone() for k, v in init.items()}
        o = AdamState(p, 1e-3, 1e-4)
        b = [idx.cpu().tolist()]
        PO.ppo_train(p, o, spec, data, b, hp)
        t0 = time.perf_counter()
        n = 10
        PO.ppo_train(p, o, spec, data, b * n, hp)
        cpu = {"value": n / (time.perf_counter() - t0), "unit": "minibatches/s", "cores": torch.get_num_threads(), "kind": "port"}
    print(json.dumps({"metric": "PPO minibatch updates/s (64 x 12x84x84 uint8, NatureCNN-512, 6 actions)",
                      "value": 1e3 / ms_graph, "unit": "minibatches/s", "us_per_minibatch_graph": ms_graph * 1e3,
                      "us_per_minibatch_eager": ms_eager * 1e3, "gpu_launches_per_minibatch": launches, "cpu_baseline": cpu}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--no-cpu", dest="cpu", action="store_false")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    from sheeprl_b200.lib import CudaOps

    cu = CudaOps()
    if args.only in ("", "sac"):
        bench_sac(args, cu)
    if args.only in ("", "ppo"):
        bench_ppo(args, cu)


if __name__ == "__main__":
    main()
