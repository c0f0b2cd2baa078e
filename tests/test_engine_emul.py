"""CPU validation of the engine's hand-written backward / kernel schedule: DV3Engine driven by the
executable op specification (oracle/ops_emul.py, a test double) must reproduce the autograd oracle and
the reference fixtures.  The same engine code drives the CUDA kernels on the GPU (tests/test_gpu_*.py)."""
import pytest
import torch

from oracle.ops_emul import EmulOps
from sheeprl_b200.engine import DV3Engine
from tests.helpers import assert_params_close, image_channels, load_fixture, oracle_run


def run_engine(cfg, adim, init, data, noise, steps, is_continuous=False):
    eng = DV3Engine(cfg, adim, in_channels=image_channels(cfg), device="cpu", ops=EmulOps(), is_continuous=is_continuous)
    eng.wm.load(init["wm"]), eng.actor.load(init["actor"]), eng.critic.load(init["critic"])
    eng.target.load(init["target"])
    outs, grads = [], []
    for s in range(steps):
        batch = {k: v.clone().float() for k, v in data[s].items()}
        eng.train_step(batch, noise[s])
        outs.append({k: float(v) for k, v in eng.metrics_dict().items()})
        grads.append({"wm": {k: v.clone() for k, v in eng.wm.gviews.items()},
                      "actor": {k: v.clone() for k, v in eng.actor.gviews.items()},
                      "critic": {k: v.clone() for k, v in eng.critic.gviews.items()}})
    return eng, outs, grads


@pytest.mark.parametrize("name", ["dv3_tiny_a", "dv3_tiny_b", "dv3_tiny_c", "dv3_tiny_v", "dv3_tiny_vo", "dv3_tiny_mk", "dv3_tiny_h0"])
def test_engine_matches_oracle_and_reference(name):
    fx, cfg = load_fixture(name)
    adim = fx["actions_dim"]
    steps = len(fx["data"])
    cont = fx.get("is_continuous", False)
    fdata = [{k: v.float() for k, v in d.items()} for d in fx["data"]]
    st, o_outs, ms, _ = oracle_run(cfg, adim, fx["init"], fdata, fx["noise"], steps, keep=True, is_continuous=cont)
    eng, e_outs, e_grads = run_engine(cfg, adim, fx["init"], fx["data"], fx["noise"], steps, is_continuous=cont)
    # gradients of the first step (pre-clip in the engine, post-clip in the oracle -> rescale)
    for grp, max_norm in (("wm", cfg.algo.world_model.clip_gradients), ("actor", cfg.algo.actor.clip_gradients),
                          ("critic", cfg.algo.critic.clip_gradients)):
        og = o_outs[0][f"grads/{grp}"]
        norm = float(o_outs[0]["Grads/" + {"wm": "world_model"}.get(grp, grp)])
        coef = min(1.0, max_norm / (norm + 1e-6))
        gmax = max(float(v.abs().max()) for v in og.values()) / coef
        gnorm = float(torch.sqrt(sum((v.double() ** 2).sum() for v in og.values())))
        for k, v in og.items():
            diff = e_grads[0][grp][k] * coef - v
            d = float(diff.abs().max())
            assert d <= 2e-5 * max(gmax * coef, 1e-12) + 1e-9, (grp, k, d, gmax)
            # every tensor on its own scale (l2), floor 1e-6 of the group norm
            rel = float(diff.double().norm()) / (float(v.double().norm()) + 1e-6 * gnorm + 1e-30)
            assert rel <= 1e-4, (grp, k, "per-tensor relative gradient error", rel)
    for s in range(steps):
        for k, v in fx["metrics"][s].items():
            assert e_outs[s][k] == pytest.approx(v, rel=3e-5, abs=1e-6), (s, k)
    lrs = {"wm": 1e-4, "actor": 8e-5, "critic": 8e-5}
    for n, g in (("wm", eng.wm), ("actor", eng.actor), ("critic", eng.critic)):
        assert_params_close(g.views, fx["after"][n], lrs[n], steps, tol=2e-6, label=n)
    assert float(eng.moments_state[0]) == pytest.approx(float(fx["moments"]["low"]), rel=1e-4, abs=1e-7)
    assert float(eng.moments_state[1]) == pytest.approx(float(fx["moments"]["high"]), rel=1e-4, abs=1e-7)


@pytest.mark.parametrize("name", ["dv3_tiny_v", "dv3_tiny_vo", "dv3_tiny_mk"])
def test_public_api_with_vector_observations(name):
    """build_agent() takes the vector dimensions from the observation space, exposes the reference's state-dict keys
    (encoder.mlp_encoder.*, observation_model.mlp_decoder.*; no cnn_* entries without an image key) and train() on a
    batch dict with one tensor per observation key lands on the executed reference"""
    from sheeprl_b200.algos.dreamer_v3.agent import build_agent
    from sheeprl_b200.algos.dreamer_v3.dreamer_v3 import make_optimizers, train
    from sheeprl_b200.algos.dreamer_v3.utils import Moments

    fx, cfg = load_fixture(name)

    class Fab:
        device = torch.device("cpu")

    class Space:
        def __init__(self, shape):
            self.shape = shape

    space = {k: Space((cfg.env.cnn_channels[k], 64, 64)) for k in cfg.algo.cnn_keys.encoder}
    space.update({k: Space((d,)) for k, d in cfg.env.mlp_dims.items()})
    cfg.env.pop("mlp_dims"), cfg.env.pop("cnn_channels")        # the product must not depend on the test-only entries
    wm0, *_ = build_agent(Fab, fx["actions_dim"], False, cfg, space, ops=EmulOps())          # default initialisation
    assert set(wm0.state_dict()) == set(fx["init"]["wm"])
    wm, actor, critic, target, player = build_agent(Fab, fx["actions_dim"], False, cfg, space, fx["init"]["wm"],
                                                    fx["init"]["actor"], fx["init"]["critic"], fx["init"]["target"],
                                                    ops=EmulOps())
    eng = wm._b200_engine
    opts = make_optimizers(eng, cfg)
    mo = cfg.algo.actor.moments
    moments = Moments(mo.decay, mo.max, mo.percentile.low, mo.percentile.high)

    class Agg:
        disabled = False
        v = {}

        def update(self, k, val):
            self.v[k] = float(val)

    for s in range(len(fx["data"])):
        agg = Agg()
        train(Fab, wm, actor, critic, target, *opts, {k: v.clone().float() for k, v in fx["data"][s].items()}, agg, cfg, False,
              fx["actions_dim"], moments, noise=fx["noise"][s])
        for k, v in fx["metrics"][s].items():
            assert agg.v[k] == pytest.approx(v, rel=3e-5, abs=1e-6), (s, k)
    assert_params_close(wm.state_dict(), fx["after"]["wm"], 1e-4, len(fx["data"]), tol=2e-6, label="wm")
