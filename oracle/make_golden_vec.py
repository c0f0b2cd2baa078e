"""TEST INFRASTRUCTURE — writes tests/golden/dv3_tiny_{v,vo}.pt by EXECUTING THE REAL REFERENCE `dreamer_v3.train`
with vector observations (container only):

    python -m oracle.make_golden_vec

dv3_tiny_mk: two image keys + one vector key;  dv3_tiny_v : one image key + two vector keys (MultiEncoder concatenates cnn and mlp features; MLPDecoder with one head
             per key, SymlogDistribution loss);  dv3_tiny_vo : vector observations only (no CNN encoder / decoder).
Same content as the fixtures of oracle/make_golden.py.
"""
from __future__ import annotations

import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle.make_golden import GOLDEN, build_case  # noqa: E402

BASE = dict(size="S", per_rank_batch_size=3, per_rank_sequence_length=5, horizon=4, dense_units=32, mlp_layers=2,
            cnn_channels_multiplier=4, recurrent_state_size=24, hidden_size=32, stochastic_size=6, discrete_size=5, bins=31,
            algo__world_model__kl_free_nats=0.05)
FIXTURES = {
    "dv3_tiny_v": dict(cfg=dict(BASE, mlp_keys={"state": 5, "extra": 3}), actions_dim=(3, 2), perturb=0.05, steps=2),
    "dv3_tiny_vo": dict(cfg=dict(BASE, cnn_keys=(), mlp_keys={"state": 7}), actions_dim=(4,), perturb=0.05, steps=2),
    # two image keys (concatenated on the channel axis by the encoder, split again by the decoder) + one vector key
    "dv3_tiny_mk": dict(cfg=dict(BASE, cnn_keys=("rgb", "depth"), cnn_channels={"rgb": 3, "depth": 1}, mlp_keys={"state": 4}),
                        actions_dim=(3,), perturb=0.05, steps=2),
    # non-default switches: the initial recurrent state is a buffer (not trained), no uniform mix
    "dv3_tiny_h0": dict(cfg=dict(BASE, algo__world_model__learnable_initial_recurrent_state=False, algo__unimix=0.0,
                                 algo__actor__unimix=0.0), actions_dim=(3,), perturb=0.05, steps=2),
}


def main():
    only = sys.argv[1:]
    for name, spec in FIXTURES.items():
        if only and name not in only:
            continue
        cfg, adim, sd, data, noise, after, metrics, moments, _ = build_case(spec)
        for d in data:
            for k in cfg.algo.cnn_keys.encoder:
                d[k] = d[k].to(torch.uint8)
        torch.save({"cfg_kwargs": spec["cfg"], "actions_dim": adim, "is_continuous": False, "init": sd, "data": data,
                    "noise": noise, "after": after, "metrics": metrics, "moments": moments}, os.path.join(GOLDEN, name + ".pt"))
        print("wrote", name, {k: round(v, 5) for k, v in metrics[-1].items()})


if __name__ == "__main__":
    main()
