"""CPU validation of the engine's hand-written backward / kernel schedule: DV3Engine driven by the
executable op specification (oracle/ops_emul.py, a test double) must reproduce the autograd oracle and
the reference fixtures.  The same engine code drives the CUDA kernels on the GPU (tests/test_gpu_*.py)."""
import pytest
import torch

from oracle.ops_emul import EmulOps
from sheeprl_b200.engine import DV3Engine
from tests.helpers import assert_params_close, load_fixture, oracle_run


def run_engine(cfg, adim, init, data, noise, steps, is_continuous=False):
    eng = DV3Engine(cfg, adim, in_channels=3, device="cpu", ops=EmulOps(), is_continuous=is_continuous)
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


@pytest.mark.parametrize("name", ["dv3_tiny_a", "dv3_tiny_b", "dv3_tiny_c"])
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
        for k, v in og.items():
            d = float((e_grads[0][grp][k] * coef - v).abs().max())
            assert d <= 2e-5 * max(gmax * coef, 1e-12) + 1e-9, (grp, k, d, gmax)
    for s in range(steps):
        for k, v in fx["metrics"][s].items():
            assert e_outs[s][k] == pytest.approx(v, rel=3e-5, abs=1e-6), (s, k)
    lrs = {"wm": 1e-4, "actor": 8e-5, "critic": 8e-5}
    for n, g in (("wm", eng.wm), ("actor", eng.actor), ("critic", eng.critic)):
        assert_params_close(g.views, fx["after"][n], lrs[n], steps, tol=2e-6, label=n)
    assert float(eng.moments_state[0]) == pytest.approx(float(fx["moments"]["low"]), rel=1e-4, abs=1e-7)
    assert float(eng.moments_state[1]) == pytest.approx(float(fx["moments"]["high"]), rel=1e-4, abs=1e-7)
