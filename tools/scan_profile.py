"""Per-phase cycle counters of the persistent RSSM scan kernels (csrc/rssm_scan.cu) at the BASELINE config:
one Dreamer-V3 update on synthetic data, counters of CTA 0 (a row owner) and CTA 1 read after each scan launch.

    python tools/scan_profile.py [--json out.json]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

FWD = {0: "prologue", 12: "B1 step start + mask-mix + sync", 13: "B1 sync + slice sums", 14: "E LayerNorm", 15: "E logits product",
       1: "B1 h-part product (thread 0's warp)", 2: "A wait z (all rows)", 3: "A gather + send x_pre, stats", 4: "B2 wait x rows + stats, normalise", 5: "B2 product+stats send",
       6: "C wait stats+merge", 7: "C gates+send h", 8: "D wait h rows", 9: "D product+send rp", 10: "E wait rp rows",
       11: "E sample + saves"}
BWD = {16: "prologue", 17: "P wait dxh_x rows", 18: "P product+softmax bwd+send", 19: "Q wait d_post_raw rows", 20: "Q product+send",
       21: "R wait dxh_r rows+sums", 22: "R product+gate bwd+send", 23: "S 3x(wait+product)+sums", 24: "S epilogue+send"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    from bench import synthetic_batch
    from sheeprl_b200.configs import make_dv3_cfg
    from sheeprl_b200.engine import DV3Engine

    cfg = make_dv3_cfg("S")
    eng = DV3Engine(cfg, (2,), in_channels=3, device="cuda:0")
    from oracle import dv3_oracle as O   # initial parameters only (tool, not product)

    wm, actor, critic, target = O.init_params(cfg, (2,), seed=0)
    eng.wm.load(wm), eng.actor.load(actor), eng.critic.load(critic), eng.target.load(target)
    data = synthetic_batch(cfg, (2,), seed=1, device="cuda:0")
    out = {}
    inner_bwd = eng.ops.rssm_scan_bwd

    def bwd(*a, **k):
        torch.cuda.synchronize()
        out["fwd"] = eng.ops.rssm_scan_profile(eng._scan_ws)
        inner_bwd(*a, **k)
        torch.cuda.synchronize()
        out["bwd"] = eng.ops.rssm_scan_profile(eng._scan_ws)

    for it in range(3):
        if it == 2:
            eng.ops.rssm_scan_bwd = bwd
        eng.train_step({k: v.clone() for k, v in data.items()})
    torch.cuda.synchronize()
    assert eng.fused_scan and eng.ops.rssm_scan_error(eng._scan_ws) == 0
    T = cfg.algo.per_rank_sequence_length
    clk_mhz = 1965.0
    rep = {}
    for name, labels in (("fwd", FWD), ("bwd", BWD)):
        for cta in (0, 1):
            tot = sum(out[name][cta])
            rows = {labels.get(i, str(i)): out[name][cta][i] for i in range(32) if out[name][cta][i]}
            rep[f"{name}_cta{cta}"] = {"total_cycles": tot, "us_per_step_at_1965MHz": tot / T / clk_mhz,
                                       "cycles_per_step": {k: round(v / T, 1) for k, v in rows.items()}}
            print(name, "cta", cta, "total cycles", tot, f"= {tot / T / clk_mhz:.2f} us/step")
            for k, v in rows.items():
                print(f"    {k:36s} {v / T:10.1f} cycles/step")
    if args.json:
        json.dump(rep, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
