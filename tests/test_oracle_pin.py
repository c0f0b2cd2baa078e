"""Pins the oracle against the LIVE reference (container only: /root/reference is not on the GPU box)."""
import pytest
import torch

from oracle import ref_harness

pytestmark = pytest.mark.skipif(not ref_harness.reference_available(), reason="reference tree absent")


def test_oracle_equals_live_reference_two_steps():
    from oracle.make_golden import FIXTURES, build_case

    spec = dict(FIXTURES["dv3_tiny_a"])
    cfg, adim, sd, data, noise, after, metrics, moments, (cp, ms) = build_case(spec, seed=3)
    from tests.helpers import assert_params_close

    for i, n in enumerate(("wm", "actor", "critic")):
        assert_params_close(cp[i], after[n], 1e-4, 2, label=n)
    assert float(ms["low"]) == pytest.approx(float(moments["low"]), rel=1e-5, abs=1e-7)


def test_reference_multinomial_is_argmax_p_over_exp():
    """torch.multinomial(p,1,True) == argmax(p / Exp(1)) on the same generator state."""
    g = torch.Generator().manual_seed(7)
    p = torch.softmax(torch.randn(64, 32, generator=g), -1)
    st = g.get_state()
    idx = torch.multinomial(p, 1, True, generator=g)
    g.set_state(st)
    q = torch.empty_like(p).exponential_(1, generator=g)
    assert torch.equal(idx.squeeze(-1), (p / q).argmax(-1))
