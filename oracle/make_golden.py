"""TEST INFRASTRUCTURE — generates tests/golden/*.pt by EXECUTING THE REAL REFERENCE (container only).

    python -m oracle.make_golden

Each fixture holds: the config kwargs, the initial state dicts (reference `build_agent` init under
torch.manual_seed, then perturbed so that every loss branch is live), the synthetic batches, the
injected Exp(1) noise (conditioned so no categorical draw sits on a near-tie) and what the unmodified
reference `train()` produced: post-step parameters, the 13 logged metrics per step, Moments buffers.
`dv3_S_digest.pt` is the BASELINE config (S, B=16, T=64, H=15): parameters are stored as a strided
subsample + per-tensor sums because the full state is 72 MB.
"""
from __future__ import annotations

import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import dv3_oracle as O  # noqa: E402
from oracle import ref_run  # noqa: E402
from sheeprl_b200.configs import make_dv3_cfg  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

FIXTURES = {
    "dv3_tiny_a": dict(
        cfg=dict(size="S", per_rank_batch_size=3, per_rank_sequence_length=5, horizon=4, dense_units=32,
                 mlp_layers=2, cnn_channels_multiplier=4, recurrent_state_size=24, hidden_size=32,
                 stochastic_size=6, discrete_size=5, bins=31, algo__world_model__kl_free_nats=0.05),
        actions_dim=(3, 2), perturb=0.05, steps=2),
    "dv3_tiny_b": dict(
        cfg=dict(size="S", per_rank_batch_size=4, per_rank_sequence_length=6, horizon=3, dense_units=48,
                 mlp_layers=1, cnn_channels_multiplier=2, recurrent_state_size=40, hidden_size=24,
                 stochastic_size=4, discrete_size=8, bins=255),
        actions_dim=(4,), perturb=0.02, steps=2),
    # continuous actions (scaled_normal): the policy gradient flows through the imagined rollout
    "dv3_tiny_c": dict(
        cfg=dict(size="S", per_rank_batch_size=3, per_rank_sequence_length=4, horizon=4, dense_units=32,
                 mlp_layers=2, cnn_channels_multiplier=4, recurrent_state_size=24, hidden_size=32,
                 stochastic_size=6, discrete_size=5, bins=31),
        actions_dim=(3,), perturb=0.05, steps=2, is_continuous=True),
}


def subsample(t: torch.Tensor, stride: int = 997) -> torch.Tensor:
    return t.flatten()[::stride].clone()


def perturbed_oracle_init(cfg, adim, seed, perturb):
    """Deterministic initial state reproducible WITHOUT the reference (used by the digest consumer)."""
    wm, actor, critic, target = O.init_params(cfg, adim, seed=seed)
    sd = {"wm": wm, "actor": actor, "critic": critic, "target": target}
    g = torch.Generator().manual_seed(5)
    for n in ("wm", "actor", "critic"):
        for v in sd[n].values():
            v.add_(torch.randn(v.shape, generator=g) * perturb)
    sd["target"] = {k: v + 0.01 for k, v in sd["critic"].items()}
    return sd


def build_case(spec, seed=0):
    cfg = make_dv3_cfg(**spec["cfg"])
    adim = tuple(spec["actions_dim"])
    cont = bool(spec.get("is_continuous", False))
    if spec.get("oracle_init"):
        sd = perturbed_oracle_init(cfg, adim, seed, spec["perturb"])
    else:
        _, _, wm, actor, critic, target, _ = ref_run.build_reference_agent(cfg, adim, seed=seed, is_continuous=cont)
        sd = ref_run.reference_state_dicts(wm, actor, critic, target)
        g = torch.Generator().manual_seed(5)
        if spec["perturb"] > 0:
            for d in sd.values():
                for v in d.values():
                    v.add_(torch.randn(v.shape, generator=g) * spec["perturb"])
            sd["target"] = {k: v + 0.01 for k, v in sd["critic"].items()}
    a, w = cfg.algo, cfg.algo.world_model
    T, B, H = a.per_rank_sequence_length, a.per_rank_batch_size, a.horizon
    steps = spec["steps"]
    data = [O.make_batch(cfg, adim, seed=1 + s, is_continuous=cont) for s in range(steps)]
    noise = [O.draw_noise(T, B, H, w.stochastic_size, w.discrete_size, adim, seed=10 + s, is_continuous=cont)
             for s in range(steps)]
    # condition the noise with the oracle (in place), then run the reference on the conditioned noise
    cp = [{k: v.clone() for k, v in sd[n].items()} for n in ("wm", "actor", "critic", "target")]
    opts = [O.AdamState(cp[0], w.optimizer.lr, w.optimizer.eps), O.AdamState(cp[1], a.actor.optimizer.lr, a.actor.optimizer.eps),
            O.AdamState(cp[2], a.critic.optimizer.lr, a.critic.optimizer.eps)]
    ms = {"low": torch.zeros(()), "high": torch.zeros(())}
    for s in range(steps):
        O.dv3_train_step(cfg, *cp, *opts, data[s], noise[s], ms, adim, condition_margin=1e-3, is_continuous=cont)
    after, metrics, moments = ref_run.run_reference_train(cfg, adim, data, noise, n_steps=steps, state=sd, seed=seed,
                                                          is_continuous=cont)
    return cfg, adim, sd, data, noise, after, metrics, moments, (cp, ms)


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    for name, spec in FIXTURES.items():
        cfg, adim, sd, data, noise, after, metrics, moments, _ = build_case(spec)
        for d in data:
            d["rgb"] = d["rgb"].to(torch.uint8)
        torch.save({"cfg_kwargs": spec["cfg"], "actions_dim": adim, "is_continuous": bool(spec.get("is_continuous", False)),
                    "init": sd, "data": data, "noise": noise,
                    "after": after, "metrics": metrics, "moments": moments}, os.path.join(GOLDEN, name + ".pt"))
        print("wrote", name, {k: round(v, 5) for k, v in metrics[-1].items()})
    # BASELINE config digest (weights are regenerated from seeds by the consumer through the oracle's
    # init, so the digest stores the reference-initialised weights' subsample for a sanity check only)
    spec = dict(cfg=dict(size="S"), actions_dim=(2,), perturb=0.02, steps=1, oracle_init=True)
    cfg, adim, sd, data, noise, after, metrics, moments, _ = build_case(spec)
    digest = {
        "cfg_kwargs": spec["cfg"], "actions_dim": adim, "data_seed": 1, "noise_seed": 10, "init_seed": 0,
        "perturb": spec["perturb"], "conditioned_noise_post_sub": subsample(noise[0]["post"]),
        "metrics": metrics, "moments": moments,
        "noise_post_sum": float(noise[0]["post"].double().sum()),
        "init_sub": {n: {k: subsample(v) for k, v in sd[n].items()} for n in sd},
        "after_sub": {n: {k: subsample(v) for k, v in after[n].items()} for n in after},
        "after_sum": {n: {k: float(v.double().sum()) for k, v in after[n].items()} for n in after},
    }
    torch.save(digest, os.path.join(GOLDEN, "dv3_S_digest.pt"))
    print("wrote dv3_S_digest", {k: round(v, 5) for k, v in metrics[-1].items()})


if __name__ == "__main__":
    main()
