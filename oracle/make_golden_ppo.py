"""TEST INFRASTRUCTURE — writes tests/golden/ppo_*.pt by EXECUTING THE REAL REFERENCE `ppo.train` (container only):

    python -m oracle.make_golden_ppo

Fixture: spec, hyper-parameters, initial parameters (reference `build_agent` under torch.manual_seed), the synthetic
rollout, the minibatch index lists the reference's RandomSampler/BatchSampler drew (recorded), the three logged
losses of every minibatch and the parameters after train().
"""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ppo_oracle as PO  # noqa: E402
from oracle import ref_harness as H  # noqa: E402
from sheeprl_b200.utils.utils import dotdict  # noqa: E402

HP = dict(clip_coef=0.2, vf_coef=1.0, ent_coef=0.0, clip_vloss=False, normalize_advantages=False, max_grad_norm=0.0)
FIXTURES = {
    # BASELINE config 0: CartPole-like vector obs, 2 discrete actions
    "ppo_vector": dict(spec=dict(cnn_channels=0, screen=0, mlp_dim=4, dense=64, layers=2, cnn_features=512,
                                 mlp_features=64, actions_dim=(2,), is_continuous=False, act="tanh"),
                       hp=dict(HP), N=96, batch=32, epochs=2, seed=11),
    # every loss branch: clipped value loss, normalised advantages, entropy bonus, gradient clipping, two heads
    "ppo_branches": dict(spec=dict(cnn_channels=0, screen=0, mlp_dim=6, dense=32, layers=2, cnn_features=512,
                                   mlp_features=16, actions_dim=(3, 2), is_continuous=False, act="tanh"),
                         hp=dict(HP, clip_vloss=True, normalize_advantages=True, ent_coef=0.01, vf_coef=0.5,
                                 max_grad_norm=0.5), N=40, batch=16, epochs=2, seed=12),
    "ppo_continuous": dict(spec=dict(cnn_channels=0, screen=0, mlp_dim=5, dense=32, layers=2, cnn_features=512,
                                     mlp_features=16, actions_dim=(3,), is_continuous=True, act="tanh"),
                           hp=dict(HP, ent_coef=0.01), N=32, batch=16, epochs=1, seed=13),
    # BASELINE config 2: pixel obs 84x84 (frame_stack 4 x rgb = 12 channels), NatureCNN, + a vector key
    # (cnn_features 48 instead of 512 keeps the committed fixture at ~2 MB; the conv stack is the full NatureCNN)
    "ppo_pixel": dict(spec=dict(cnn_channels=12, screen=84, mlp_dim=3, dense=64, layers=2, cnn_features=48,
                                mlp_features=64, actions_dim=(4,), is_continuous=False, act="tanh"),
                      hp=dict(HP, ent_coef=0.01), N=12, batch=4, epochs=1, seed=14),
    # non-default variants: tanh-squashed Normal policy, LayerNorm MLPs, ReLU, different widths/depths per network,
    # two vector keys (concatenated by the encoder)
    "ppo_tanh_ln": dict(spec=dict(cnn_channels=0, screen=0, mlp_dim=5, mlp_keys=[("state", 3), ("extra", 2)], dense=32,
                                  layers=2, nets={"encoder": (24, 1), "actor": (32, 2), "critic": (16, 3)}, layer_norm=True,
                                  cnn_features=512, mlp_features=16, actions_dim=(3,), is_continuous=True,
                                  dist="tanh_normal", act="relu"),
                        hp=dict(HP, clip_vloss=True, normalize_advantages=True, ent_coef=0.01, vf_coef=0.5,
                                max_grad_norm=0.5), N=40, batch=16, epochs=2, seed=15),
    # two image keys (channel concat) + two vector keys, LayerNorm in encoder and critic only, 64x64 screen
    "ppo_multikey": dict(spec=dict(cnn_channels=4, cnn_keys=[("rgb", 3), ("depth", 1)], screen=64, mlp_dim=5,
                                   mlp_keys=[("state", 3), ("extra", 2)], dense=32, layers=2,
                                   layer_norm={"encoder": True, "actor": False, "critic": True}, cnn_features=32,
                                   mlp_features=16, actions_dim=(3,), is_continuous=False, act="tanh"),
                         hp=dict(HP, ent_coef=0.01, max_grad_norm=0.5), N=10, batch=5, epochs=1, seed=16),
}


def obs_keys(spec):
    """[(key, channels)], [(key, dim)] of the image / vector observation keys of a fixture spec"""
    cnn = list(spec.get("cnn_keys") or ([("rgb", spec["cnn_channels"])] if spec["cnn_channels"] else []))
    mlp = list(spec.get("mlp_keys") or ([("state", spec["mlp_dim"])] if spec["mlp_dim"] else []))
    return cnn, mlp


def split_obs(spec, d):
    """the reference's per-key view of a rollout / obs dict that holds the concatenated "rgb" / "state" tensors"""
    out = {k: v for k, v in d.items() if k not in ("rgb", "state")}
    cnn, mlp = obs_keys(spec)
    off = 0
    for k, c in cnn:
        out[k] = d["rgb"][:, off:off + c].clone()
        off += c
    off = 0
    for k, c in mlp:
        out[k] = d["state"][:, off:off + c].clone()
        off += c
    return out


def ppo_cfg(spec, hp, batch, epochs):
    def net(which):
        dense, layers, ln = PO.net_cfg(spec, which)
        return {"dense_units": dense, "mlp_layers": layers, "layer_norm": ln, "ortho_init": False,
                "dense_act": "torch.nn.Tanh" if spec["act"] == "tanh" else "torch.nn.ReLU"}

    cnn, mlp = obs_keys(spec)
    enc = dict(net("encoder"), cnn_features_dim=spec["cnn_features"], mlp_features_dim=spec["mlp_features"])
    dist = spec.get("dist", "auto") if spec["is_continuous"] else "auto"
    return dotdict({"algo": dict(hp, cnn_keys={"encoder": [k for k, _ in cnn]}, mlp_keys={"encoder": [k for k, _ in mlp]},
                                 encoder=enc, actor=net("actor"), critic=net("critic"), per_rank_batch_size=batch,
                                 update_epochs=epochs, loss_reduction="mean",
                                 optimizer={"lr": 1e-3, "eps": 1e-4, "weight_decay": 0, "betas": [0.9, 0.999]}),
                    "buffer": {"share_data": False}, "env": {"screen_size": spec["screen"]}, "seed": 0,
                    "distribution": {"type": dist}})


def obs_space(spec):
    cnn, mlp = obs_keys(spec)
    space = {k: H.Shape((c, spec["screen"], spec["screen"])) for k, c in cnn}
    space.update({k: H.Shape((d,)) for k, d in mlp})
    return space


def run(fx):
    import sheeprl.algos.ppo.agent as A
    import sheeprl.algos.ppo.ppo as P

    A.get_single_device_fabric = lambda f: f
    spec, hp = fx["spec"], fx["hp"]
    cfg = ppo_cfg(spec, hp, fx["batch"], fx["epochs"])
    space = obs_space(spec)
    fab = H.FakeFabric()
    torch.manual_seed(fx["seed"])
    agent, _ = A.build_agent(fab, spec["actions_dim"], spec["is_continuous"], cfg, space, None)
    export = lambda: {k.replace('_forward_module.', ''): v.detach().clone() for k, v in agent.state_dict().items()}  # noqa: E731
    init = export()
    assert {k: tuple(v.shape) for k, v in init.items()} == PO.ppo_param_shapes(spec), "key/shape layout drifted"
    opt = torch.optim.Adam(agent.parameters(), lr=1e-3, eps=1e-4)
    data = PO.make_rollout(spec, fx["N"], fx["seed"] + 1)
    # record the minibatch indices the reference draws
    drawn = []
    orig = P.BatchSampler

    class Recording(orig):
        def __iter__(self):
            for b in super().__iter__():
                drawn.append(list(b))
                yield b

    class Agg:
        disabled = False

        def __init__(self):
            self.rows = []

        def update(self, k, v):
            if k == "Loss/policy_loss":
                self.rows.append({})
            self.rows[-1][k] = float(v)

    agg = Agg()
    P.BatchSampler = Recording
    try:
        torch.manual_seed(fx["seed"] + 2)
        P.train(fab, agent, opt, split_obs(spec, {k: v.clone() for k, v in data.items()}), agg, cfg)
    finally:
        P.BatchSampler = orig
    after = export()
    return {"spec": spec, "hp": hp, "batch": fx["batch"], "epochs": fx["epochs"], "init": init, "data": data,
            "index_batches": drawn, "losses": agg.rows, "after": after, "sampler_seed": fx["seed"] + 2}


def main():
    H.install()
    for name, fx in FIXTURES.items():
        out = run(fx)
        if "rgb" in out["data"]:                      # integer-valued pixels: keep the fixture small
            out["data"]["rgb"] = out["data"]["rgb"].to(torch.uint8)
        path = os.path.join(ROOT, "tests", "golden", f"{name}.pt")
        torch.save(out, path)
        print(name, os.path.getsize(path), len(out["index_batches"]), out["losses"][-1])


if __name__ == "__main__":
    main()
