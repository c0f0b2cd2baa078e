"""The HBM-bound kernels of the hot path, each launched on the shape it has in the Dreamer-V3 S step (or at the
replay / PPO shapes of SURVEY §8 a17-a19), for ONE `ncu --set full` capture and for an in-process timing table.

    python tools/hbm_kernels.py                      # CUDA-event timing -> one JSON line (algorithmic GB/s per kernel)
    ncu --set full --clock-control none --profile-from-start off -o gpurun_out/r1_hbm_kernels python tools/hbm_kernels.py --once

With --once every kernel runs exactly once inside a cudaProfilerStart/Stop window (after a warm-up outside it) with
the L2 flushed before each launch.  Algorithmic bytes = the tensors each kernel has to read and write once.
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def build_xl_layernorm(cu, dev="cuda"):
    """the LayerNorm launches of the XL model (BASELINE configs[4]: conv channels 96..768 on 4096 images, dense 1024 on the
    65 536 imagined rows, the GRU's joint LayerNorm over 3 x 4096)"""
    f = lambda *s: torch.randn(*s, device=dev)  # noqa: E731
    z = lambda *s: torch.zeros(*s, device=dev)  # noqa: E731
    out = []
    for M, C in ((4096 * 32 * 32, 96), (4096 * 16 * 16, 192), (4096 * 8 * 8, 384), (4096 * 4 * 4, 768), (65536, 1024),
                 (4096, 12288), (64, 12288), (64, 1024)):
        X, Y, dY, dX = f(M, C), z(M, C), f(M, C), z(M, C)
        gam, bet, dg, db = f(C) + 1, f(C), z(C), z(C)
        out.append((f"ln_act_fwd [{M},{C}]", 8 * M * C, lambda X=X, Y=Y, gam=gam, bet=bet: cu.ln_act_fwd(X, gam, bet, 1e-3, 1, Y)))
        out.append((f"ln_act_bwd [{M},{C}]", 12 * M * C,
                    lambda X=X, dY=dY, dX=dX, gam=gam, bet=bet, dg=dg, db=db: cu.ln_act_bwd(X, gam, bet, 1e-3, 1, dY, dX, dg, db)))
    return out


def build(cu, dev="cuda"):
    """[(name, algorithmic_bytes, thunk)]"""
    f = lambda *s: torch.randn(*s, device=dev)  # noqa: E731
    z = lambda *s: torch.zeros(*s, device=dev)  # noqa: E731
    out = []

    # a17: replay gather, 16 x (16 sequences x 64 rows x 12 288 B) from a 1.6 GB ring
    T, B, G = 64, 16, 16
    store = torch.randint(0, 256, (131072, 3, 64, 64), dtype=torch.uint8, device=dev)
    idx = (torch.randint(0, store.shape[0] - T, (G * B, 1), device=dev) + torch.arange(T, device=dev)).reshape(-1).contiguous()
    go = torch.empty(G, T, B, 3, 64, 64, dtype=torch.uint8, device=dev)
    out.append(("replay_gather u8 16x[64,16,3,64,64]", 2 * go.numel(), lambda: cu.replay_gather(store, idx, go, G, B, T)))

    # a1: uint8 -> fp32 normalise + NHWC
    obs = torch.randint(0, 256, (1024, 3, 64, 64), dtype=torch.uint8, device=dev)
    x0 = z(1024, 64, 64, 3)
    out.append(("obs_prep u8->f32 [1024,3,64,64]", obs.numel() + 4 * x0.numel(), lambda: cu.obs_prep(obs, x0)))

    # a2/a4: LayerNorm+SiLU of the conv stacks (first encoder stage: 1 M rows x 32 channels) and of an MLP layer
    for M, C in ((1024 * 32 * 32, 32), (1024 * 16 * 16, 64), (16384, 512)):
        X, Y, dY, dX = f(M, C), z(M, C), f(M, C), z(M, C)
        gam, bet, dg, db = f(C) + 1, f(C), z(C), z(C)
        out.append((f"ln_act_fwd [{M},{C}]", 8 * M * C, lambda X=X, Y=Y, gam=gam, bet=bet: cu.ln_act_fwd(X, gam, bet, 1e-3, 1, Y)))
        out.append((f"ln_act_bwd [{M},{C}]", 12 * M * C,
                    lambda X=X, dY=dY, dX=dX, gam=gam, bet=bet, dg=dg, db=db: cu.ln_act_bwd(X, gam, bet, 1e-3, 1, dY, dX, dg, db)))

    # a6: reconstruction MSE (value + gradient), two-hot CE, KL
    pred, tgt, grad, rowl = f(1024, 12288), f(1024, 12288), z(1024, 12288), z(1024)
    out.append(("mse_loss_grad [1024,12288]", 12 * pred.numel(), lambda: cu.mse_loss_grad(pred, tgt, 1.0 / 1024, rowl, grad)))
    lg, xs, wt, lr_, dl = f(15360, 255), f(15360), torch.ones(15360, device=dev), z(15360), z(15360, 255)
    out.append(("twohot_loss_grad [15360,255]", 8 * lg.numel(), lambda: cu.twohot_loss_grad(lg, xs, wt, 1.0 / 15360, -20.0, 20.0, lr_, dl)))
    lg2, mean = f(16384, 255), z(16384)
    out.append(("twohot_mean [16384,255]", 4 * lg2.numel(), lambda: cu.twohot_mean(lg2, -20.0, 20.0, mean)))
    pm, qm, dp, dq, rows = f(1024, 1024), f(1024, 1024), z(1024, 1024), z(1024, 1024), z(1024, 4)
    out.append(("kl_loss_grad [1024,32x32]", 16 * pm.numel(),
                lambda: cu.kl_loss_grad(pm, qm, 32, 32, 0.5, 0.1, 1.0, 1.0, 1.0 / 1024, dp, dq, rows)))

    # a7: global norm + fused clip+Adam on the world model's flat group (15.69 M parameters, 28 B / parameter)
    n = 15_690_000
    p, g, m, v = f(n), f(n) * 1e-3, z(n), z(n)
    nsq, step, nout = torch.zeros(1, dtype=torch.float64, device=dev), torch.ones(1, dtype=torch.int32, device=dev), z(1)
    out.append(("sumsq [15.69M]", 4 * n, lambda: cu.sumsq(g, nsq)))
    out.append(("adam_step [15.69M]", 28 * n, lambda: cu.adam_step(p, g, m, v, nsq, 1000.0, 1e-4, 0.9, 0.999, 1e-8, step, nout)))
    # a14: target EMA
    tc, sc = f(1_180_000), f(1_180_000)
    out.append(("ema [1.18M]", 12 * tc.numel(), lambda: cu.ema(tc, sc, 0.02)))

    # a10 / a18: the reverse scans (latency-sized)
    H, N = 15, 1024
    rew, val, cl, tcnt, lam, disc = f(H + 1, N), f(H + 1, N), f(H + 1, N), torch.ones(N, device=dev), z(H, N), z(H + 1, N)
    out.append(("lambda_returns [16,1024]", 4 * 6 * (H + 1) * N, lambda: cu.lambda_returns(rew, val, cl, tcnt, 0.997, 0.95, lam, disc)))
    r, vv, d, nv = f(128, 16, 1), f(128, 16, 1), (f(128, 16, 1) > 2).float(), f(16, 1)
    ret, adv = torch.empty_like(r), torch.empty_like(r)
    out.append(("gae [128,16]", 4 * 5 * 128 * 16, lambda: cu.gae(r, vv, d, nv, 0.99, 0.95, ret, adv)))
    return out


def main():
    from sheeprl_b200.lib import CudaOps

    cu = CudaOps()
    kernels = build_xl_layernorm(cu) if "--xl-layernorm" in sys.argv else build(cu)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")          # > 126 MB L2
    for _, _, fn in kernels:                                                   # warm-up (module load, scratch)
        fn()
    torch.cuda.synchronize()
    if "--once" in sys.argv:
        for _, _, fn in kernels:
            flush.zero_()
            torch.cuda.synchronize()
            torch.cuda.profiler.start()
            fn()
            torch.cuda.synchronize()
            torch.cuda.profiler.stop()
        return
    table = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for name, nbytes, fn in kernels:
        ms = 0.0
        for _ in range(10):
            flush.zero_()
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ms += e0.elapsed_time(e1)
        ms /= 10
        table[name] = {"us": round(ms * 1e3, 2), "algorithmic_MB": round(nbytes / 1e6, 2), "GBps": round(nbytes / ms / 1e6, 1)}
    print(json.dumps({"hbm_kernels": table}))


if __name__ == "__main__":
    main()
