"""HBM kernels outside the train step: replay gather (a17) at the Dreamer-V3 batch (16 sequences x 64 rows x 12 288 B,
uint8) from a 4-env x 65 536-row device ring (3.2 GB), through the public buffer API, and the PPO GAE scan (a18,
[128, 16]).  Prints one JSON line; algorithmic bytes = 2 x gathered bytes (read + write).

    python tools/bench_replay.py
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from sheeprl_b200.data.buffers import EnvIndependentReplayBuffer, SequentialReplayBuffer
    from sheeprl_b200.lib import CudaOps

    cu = CudaOps()
    size, n_envs, T, B = 65536, 4, 64, 16
    rb = EnvIndependentReplayBuffer(size, n_envs, buffer_cls=SequentialReplayBuffer, device="cuda", ops=cu)
    g = np.random.default_rng(0)
    chunk = 8192
    rows = {"rgb": g.integers(0, 256, size=(chunk, n_envs, 3, 64, 64), dtype=np.uint8),
            "rewards": g.standard_normal((chunk, n_envs, 1))}
    for _ in range(size // chunk + 1):                          # wraps every ring once
        rb.add(rows)
    out = {}
    for n_samples in (1, 16):
        for _ in range(3):
            rb.sample_tensors(B, n_samples=n_samples, sequence_length=T)
        # kernel-only: reuse one index plan
        store = rb._flat("rgb")
        idx = torch.randint(0, store.shape[0] - T, (n_samples * B, 1), device="cuda") + torch.arange(T, device="cuda")
        idx = idx.reshape(-1).contiguous()
        o = torch.empty(n_samples, T, B, 3, 64, 64, dtype=torch.uint8, device="cuda")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        ms = 0.0
        reps = 20
        for _ in range(reps):
            flush.zero_()                                        # evict L2 between launches
            e0.record()
            cu.replay_gather(store, idx, o, n_samples, B, T)
            e1.record()
            torch.cuda.synchronize()
            ms += e0.elapsed_time(e1)
        ms /= reps
        nbytes = 2 * o.numel()
        # public API (host index plan + H2D of indices + one launch per key), steady state
        e0.record()
        for _ in range(reps):
            rb.sample_tensors(B, n_samples=n_samples, sequence_length=T)
        e1.record()
        torch.cuda.synchronize()
        out[f"n_samples={n_samples}"] = {"gather_us": ms * 1e3, "algorithmic_GBps": nbytes / ms / 1e6,
                                         "bytes": nbytes, "public_api_us": e0.elapsed_time(e1) / reps * 1e3}
    Tn, E = 128, 16
    r, v, d, nv = (torch.randn(Tn, E, 1, device="cuda") for _ in range(3)), None, None, None
    r, v, d = r
    d = (d > 2).float()
    nv = torch.randn(E, 1, device="cuda")
    ret, adv = torch.empty_like(r), torch.empty_like(r)
    for _ in range(3):
        cu.gae(r, v, d, nv, 0.99, 0.95, ret, adv)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(100):
        cu.gae(r, v, d, nv, 0.99, 0.95, ret, adv)
    e1.record()
    torch.cuda.synchronize()
    out["gae_128x16_us"] = e0.elapsed_time(e1) * 10
    print(json.dumps({"replay_gather": out}))


if __name__ == "__main__":
    main()
