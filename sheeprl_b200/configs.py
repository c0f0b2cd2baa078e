"""Resolved Dreamer-V3 configuration trees (host side).

The reference composes these with Hydra from `sheeprl/configs/algo/dreamer_v3*.yaml`,
`configs/optim/adam.yaml` and `configs/exp/dreamer_v3.yaml`; after composition `train()` only ever
sees a nested attribute dict (reference: sheeprl/cli.py:364).  This module builds the same resolved
tree directly (values: SURVEY.md Appendix D; reference: configs/algo/dreamer_v3.yaml:1-165,
configs/algo/dreamer_v3_{XS,S,M,L,XL}.yaml, configs/exp/dreamer_v3.yaml:10-26).
"""
from __future__ import annotations

import copy
from typing import Any, Dict, Mapping, Sequence

from sheeprl_b200.utils.utils import dotdict

# size -> (dense_units, mlp_layers, cnn_multiplier, recurrent_state_size, hidden_size)
DV3_SIZES: Dict[str, Sequence[int]] = {
    "XS": (256, 1, 24, 256, 256),
    "S": (512, 2, 32, 512, 512),
    "M": (640, 3, 48, 1024, 640),
    "L": (768, 4, 64, 2048, 768),
    "XL": (1024, 5, 96, 4096, 1024),
}


def _adam(lr: float, eps: float) -> Dict[str, Any]:
    return {"_target_": "torch.optim.Adam", "lr": lr, "eps": eps, "weight_decay": 0, "betas": [0.9, 0.999]}


def make_dv3_cfg(
    size: str = "S",
    *,
    per_rank_batch_size: int = 16,
    per_rank_sequence_length: int = 64,
    horizon: int = 15,
    screen_size: int = 64,
    num_envs: int = 4,
    cnn_keys: Sequence[str] = ("rgb",),
    mlp_keys: Mapping[str, int] | None = None,
    cnn_channels: Mapping[str, int] | None = None,
    dense_units: int | None = None,
    mlp_layers: int | None = None,
    cnn_channels_multiplier: int | None = None,
    recurrent_state_size: int | None = None,
    hidden_size: int | None = None,
    stochastic_size: int = 32,
    discrete_size: int = 32,
    bins: int = 255,
    **overrides: Any,
) -> dotdict:
    """Build the resolved `cfg` tree `train()`/`build_agent()` read. Keyword overrides named like
    `algo.world_model.kl_free_nats=0.1` may be passed through `overrides` with `__` for dots.
    mlp_keys: {vector observation key: dimension}; the dimensions (which the reference reads from the observation
    space) are kept under `env.mlp_dims` for the synthetic batches / initialisers."""
    mlp_keys = dict(mlp_keys or {})
    cnn_channels = {k: int((cnn_channels or {}).get(k, 3)) for k in cnn_keys}     # channels per image key (synthetic data)
    du, ml, mult, rss, hs = DV3_SIZES[size]
    du = dense_units or du
    ml = mlp_layers or ml
    mult = cnn_channels_multiplier or mult
    rss = recurrent_state_size or rss
    hs = hidden_size or hs
    mlp_ln = {"cls": "sheeprl.models.models.LayerNorm", "kw": {"eps": 1e-3}}
    cnn_ln = {"cls": "sheeprl.models.models.LayerNormChannelLast", "kw": {"eps": 1e-3}}
    act = "torch.nn.SiLU"

    def head(extra=None):
        d = {"dense_act": act, "mlp_layers": ml, "layer_norm": copy.deepcopy(mlp_ln), "dense_units": du}
        d.update(extra or {})
        return d

    cfg = {
        "seed": 42,
        "dry_run": False,
        "env": {"screen_size": screen_size, "num_envs": num_envs, "mlp_dims": mlp_keys, "cnn_channels": cnn_channels},
        "distribution": {"type": "auto", "validate_args": False},
        "algo": {
            "name": "dreamer_v3",
            "gamma": 0.996996996996997,
            "lmbda": 0.95,
            "horizon": horizon,
            "replay_ratio": 1,
            "learning_starts": 1024,
            "per_rank_pretrain_steps": 0,
            "per_rank_batch_size": per_rank_batch_size,
            "per_rank_sequence_length": per_rank_sequence_length,
            "total_steps": 5000000,
            "run_test": True,
            "cnn_keys": {"encoder": list(cnn_keys), "decoder": list(cnn_keys)},
            "mlp_keys": {"encoder": list(mlp_keys), "decoder": list(mlp_keys)},
            "cnn_layer_norm": copy.deepcopy(cnn_ln),
            "mlp_layer_norm": copy.deepcopy(mlp_ln),
            "dense_units": du,
            "mlp_layers": ml,
            "dense_act": act,
            "cnn_act": act,
            "unimix": 0.01,
            "hafner_initialization": True,
            "world_model": {
                "discrete_size": discrete_size,
                "stochastic_size": stochastic_size,
                "kl_dynamic": 0.5,
                "kl_representation": 0.1,
                "kl_free_nats": 1.0,
                "kl_regularizer": 1.0,
                "continue_scale_factor": 1.0,
                "clip_gradients": 1000.0,
                "decoupled_rssm": False,
                "learnable_initial_recurrent_state": True,
                "encoder": {
                    "cnn_channels_multiplier": mult, "cnn_act": act, "dense_act": act, "mlp_layers": ml,
                    "cnn_layer_norm": copy.deepcopy(cnn_ln), "mlp_layer_norm": copy.deepcopy(mlp_ln),
                    "dense_units": du,
                },
                "recurrent_model": {"recurrent_state_size": rss, "layer_norm": copy.deepcopy(mlp_ln),
                                    "dense_units": du},
                "transition_model": {"hidden_size": hs, "dense_act": act, "layer_norm": copy.deepcopy(mlp_ln)},
                "representation_model": {"hidden_size": hs, "dense_act": act,
                                         "layer_norm": copy.deepcopy(mlp_ln)},
                "observation_model": {
                    "cnn_channels_multiplier": mult, "cnn_act": act, "dense_act": act, "mlp_layers": ml,
                    "cnn_layer_norm": copy.deepcopy(cnn_ln), "mlp_layer_norm": copy.deepcopy(mlp_ln),
                    "dense_units": du,
                },
                "reward_model": head({"bins": bins}),
                "discount_model": head({"learnable": True}),
                "optimizer": _adam(1e-4, 1e-8),
            },
            "actor": {
                "cls": "sheeprl.algos.dreamer_v3.agent.Actor",
                "ent_coef": 3e-4, "min_std": 0.1, "max_std": 1.0, "init_std": 2.0,
                "dense_act": act, "mlp_layers": ml, "layer_norm": copy.deepcopy(mlp_ln), "dense_units": du,
                "clip_gradients": 100.0, "unimix": 0.01, "action_clip": 1.0,
                "moments": {"decay": 0.99, "max": 1.0, "percentile": {"low": 0.05, "high": 0.95}},
                "optimizer": _adam(8e-5, 1e-5),
            },
            "critic": {
                "dense_act": act, "mlp_layers": ml, "layer_norm": copy.deepcopy(mlp_ln), "dense_units": du,
                "per_rank_target_network_update_freq": 1, "tau": 0.02, "bins": bins, "clip_gradients": 100.0,
                "optimizer": _adam(8e-5, 1e-5),
            },
            "player": {"discrete_size": discrete_size},
        },
    }
    cfg = dotdict(cfg)
    for k, v in overrides.items():
        node = cfg
        parts = k.split("__")
        for p in parts[:-1]:
            node = node[p]
        node[parts[-1]] = v
    return cfg


def make_p2e_dv3_cfg(size: str = "S", *, n_ensembles: int = 8, intrinsic_weight: float = 0.1, extrinsic_weight: float = 1.0,
                     intrinsic_reward_multiplier: float = 1.0, **kw: Any) -> dotdict:
    """`make_dv3_cfg` plus the Plan2Explore additions of `configs/algo/p2e_dv3.yaml` (ensembles, the exploration
    critics and their mixing weights, the player's actor choice)."""
    cfg = make_dv3_cfg(size, **kw)
    a = cfg.algo
    a["name"] = "p2e_dv3_exploration"
    a["intrinsic_reward_multiplier"] = intrinsic_reward_multiplier
    a["player"]["actor_type"] = "exploration"
    a["critics_exploration"] = dotdict({"intrinsic": {"weight": intrinsic_weight, "reward_type": "intrinsic"},
                                        "extrinsic": {"weight": extrinsic_weight, "reward_type": "task"}})
    a["ensembles"] = dotdict({"n": n_ensembles, "dense_act": a.dense_act, "mlp_layers": a.mlp_layers,
                              "dense_units": a.dense_units, "layer_norm": copy.deepcopy(a.mlp_layer_norm.as_dict()),
                              "clip_gradients": 100.0, "optimizer": _adam(1e-4, 1e-5)})
    return cfg
