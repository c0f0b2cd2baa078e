"""The oracle (oracle/dv3_oracle.py) must reproduce what the EXECUTED reference produced
(fixtures written by oracle/make_golden.py from the unmodified reference train())."""
import pytest
import torch

from tests.helpers import assert_params_close, load_fixture, oracle_run

METRICS_RTOL = 2e-5


@pytest.mark.parametrize("name", ["dv3_tiny_a", "dv3_tiny_b", "dv3_tiny_c", "dv3_tiny_v", "dv3_tiny_vo", "dv3_tiny_mk", "dv3_tiny_h0"])
def test_oracle_matches_reference_fixture(name):
    fx, cfg = load_fixture(name)
    steps = len(fx["data"])
    st, outs, ms, _ = oracle_run(cfg, fx["actions_dim"], fx["init"], [{k: v.float() for k, v in d.items()} for d in fx["data"]],
                                 fx["noise"], steps, is_continuous=fx.get("is_continuous", False))
    for s in range(steps):
        for k, v in fx["metrics"][s].items():
            assert float(outs[s][k]) == pytest.approx(v, rel=METRICS_RTOL, abs=1e-6), (s, k)
    lrs = {"wm": cfg.algo.world_model.optimizer.lr, "actor": cfg.algo.actor.optimizer.lr,
           "critic": cfg.algo.critic.optimizer.lr}
    for n in ("wm", "actor", "critic"):
        assert_params_close(st[n], fx["after"][n], lrs[n], steps, label=n)
    assert float(ms["low"]) == pytest.approx(float(fx["moments"]["low"]), rel=1e-5, abs=1e-7)
    assert float(ms["high"]) == pytest.approx(float(fx["moments"]["high"]), rel=1e-5, abs=1e-7)


def test_oracle_matches_reference_baseline_config_digest():
    """BASELINE config (S, B=16, T=64, H=15): metrics + strided parameter subsample of the reference."""
    from oracle import dv3_oracle as O
    from oracle.make_golden import perturbed_oracle_init, subsample

    fx, cfg = load_fixture("dv3_S_digest")
    adim = fx["actions_dim"]
    a, w = cfg.algo, cfg.algo.world_model
    init = perturbed_oracle_init(cfg, adim, fx["init_seed"], fx["perturb"])
    for n in init:
        for k, v in init[n].items():
            assert torch.equal(subsample(v), fx["init_sub"][n][k]), (n, k)
    data = [O.make_batch(cfg, adim, seed=fx["data_seed"])]
    noise = [O.draw_noise(a.per_rank_sequence_length, a.per_rank_batch_size, a.horizon, w.stochastic_size,
                          w.discrete_size, adim, seed=fx["noise_seed"])]
    st, outs, ms, _ = oracle_run(cfg, adim, init, data, noise, 1, condition_margin=1e-3)
    for k, v in fx["metrics"][0].items():
        assert float(outs[0][k]) == pytest.approx(v, rel=1e-4, abs=1e-6), k
    for n in ("wm", "actor", "critic"):
        sub = {k: subsample(v) for k, v in st[n].items()}
        assert_params_close(sub, fx["after_sub"][n], 1e-4, 1, tol=2e-6, frac=5e-3, label=n)
