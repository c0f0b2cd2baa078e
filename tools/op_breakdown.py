"""Per-(op, shape) timing of one eager Dreamer-V3 update at the BASELINE config (CUDA events around every C-ABI call):
which products / convolutions / LayerNorms the step time is made of.

    python tools/op_breakdown.py [--size S|XL] [--top 40] [--json out.json]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


class ShapeProfiler:
    def __init__(self, inner):
        self._inner, self.records = inner, []

    def __getattr__(self, name):
        fn = getattr(self._inner, name)
        if not callable(fn):
            return fn

        def wrapped(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            shapes = ",".join("x".join(map(str, t.shape)) for t in a[:4] if hasattr(t, "shape"))
            flags = ",".join(str(x) for x in a if isinstance(x, bool))
            e0.record()
            r = fn(*a, **k)
            e1.record()
            self.records.append((f"{name}[{shapes}]{('{' + flags + '}') if flags else ''}", e0, e1))
            return r

        return wrapped


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", default="S")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    from bench import synthetic_batch
    from sheeprl_b200.algos.dreamer_v3.agent import initial_state
    from sheeprl_b200.configs import make_dv3_cfg
    from sheeprl_b200.engine import DV3Engine

    cfg = make_dv3_cfg(args.size, per_rank_batch_size=16 if args.size == "S" else 64)
    eng = DV3Engine(cfg, (2,), device="cuda")
    g = torch.Generator().manual_seed(0)
    for grp in (eng.wm, eng.actor, eng.critic):
        grp.load(initial_state(grp, {}, g))
    eng.target.load(eng.critic.state_dict())
    data = synthetic_batch(cfg, (2,), 1, device="cuda")
    for _ in range(3):
        eng.train_step({k: v.clone() for k, v in data.items()}, None)
    prof = ShapeProfiler(eng.ops)
    eng.ops = prof
    for m in (eng.reward_wm, eng.cont_wm, eng.actor_mlp, eng.critic_mlp, eng.target_mlp, eng.rew_img, eng.cont_img):
        m.eng = eng
    eng.train_step({k: v.clone() for k, v in data.items()}, None)
    torch.cuda.synchronize()
    agg = {}
    for key, e0, e1 in prof.records:
        s = agg.setdefault(key, [0.0, 0])
        s[0] += e0.elapsed_time(e1)
        s[1] += 1
    tot = sum(v[0] for v in agg.values())
    rows = sorted(agg.items(), key=lambda kv: -kv[1][0])
    print(f"total {tot:.2f} ms over {len(prof.records)} calls")
    for k, (ms, n) in rows[: args.top]:
        print(f"{ms:8.3f} ms  {100 * ms / tot:5.1f} %  x{n:<4d} {1e3 * ms / n:8.1f} us/call  {k}")
    if args.json:
        json.dump({"total_ms": tot, "rows": [{"op": k, "ms": ms, "calls": n} for k, (ms, n) in rows]}, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
