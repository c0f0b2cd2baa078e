"""Known-answer vectors of the reference's own two-hot tests (tests/test_utils/test_two_hot_encoder.py:6-88, SURVEY §8c:
2.3 over [-5, 5] with 11 buckets -> {0.7 @ 7, 0.3 @ 8}; 21 buckets -> {0.4 @ 14, 0.6 @ 15}; +-6.1 saturate; integers and
the corners are exact) applied to the two-hot target of `b200rl_twohot_loss_grad`.  That kernel encodes symlog(x)
(TwoHotEncodingDistribution.log_prob, utils/distribution.py:253-276), so x = symexp(v) is fed for a support value v; with
zero logits the softmax is uniform and  target = 1/nb - dlogits."""
import math

import pytest
import torch

KATS = [  # (support value, low, high, buckets, {index: weight})
    (2.3, -5.0, 5.0, 11, {7: 0.7, 8: 0.3}),
    (2.3, -5.0, 5.0, 21, {14: 0.4, 15: 0.6}),
    (3.4, -5.0, 5.0, 11, {8: 0.6, 9: 0.4}),
    (6.1, -5.0, 5.0, 11, {10: 1.0}),
    (-6.1, -5.0, 5.0, 11, {0: 1.0}),
    (2.0, -5.0, 5.0, 11, {7: 1.0}),
    (5.0, -5.0, 5.0, 11, {10: 1.0}),
    (-5.0, -5.0, 5.0, 11, {0: 1.0}),
]


def symexp(v: float) -> float:
    return math.copysign(math.exp(abs(v)) - 1.0, v)


def twohot_targets(ops, device="cpu"):
    out = []
    for v, low, high, nb, _ in KATS:
        logits = torch.zeros(1, nb, device=device)
        x = torch.tensor([symexp(v)], dtype=torch.float32, device=device)
        rows, dl = torch.zeros(1, device=device), torch.zeros(1, nb, device=device)
        ops.twohot_loss_grad(logits, x, None, 1.0, low, high, rows, dl)
        out.append((1.0 / nb - dl[0]).cpu())
    return out


def check(targets):
    for (v, low, high, nb, want), got in zip(KATS, targets):
        exp = torch.zeros(nb)
        for i, w in want.items():
            exp[i] = w
        assert torch.allclose(got, exp, atol=2e-5), (v, nb, got.tolist())
        assert float(got.sum()) == pytest.approx(1.0, abs=1e-5)


def test_op_specification_reproduces_the_reference_vectors():
    from oracle.ops_emul import EmulOps

    check(twohot_targets(EmulOps()))


def test_oracle_log_prob_reproduces_the_reference_vectors():
    """-log_prob with zero logits = log(nb) for any target that sums to one; the gradient w.r.t. the logits is
    softmax - target"""
    from oracle.dv3_oracle import twohot_log_prob

    for v, low, high, nb, want in KATS:
        logits = torch.zeros(1, nb, requires_grad=True)
        lp = twohot_log_prob(logits, torch.tensor([[symexp(v)]]), low, high)
        (-lp.sum()).backward()
        tgt = 1.0 / nb - logits.grad[0]
        exp = torch.zeros(nb)
        for i, w in want.items():
            exp[i] = w
        assert torch.allclose(tgt, exp, atol=2e-5), (v, nb, tgt.tolist())


def test_decoder_vectors_on_the_mean_op():
    """tests/test_utils/test_two_hot_decoder.py:6-60: {0.7 @ 7, 0.3 @ 8} -> 2.3, {0.6 @ 8, 0.4 @ 9} -> 3.4, one-hot bins ->
    their value.  `twohot_mean` returns symexp of the expected bin (TwoHotEncodingDistribution.mean,
    utils/distribution.py:245-247), so the expectation is symexp(KAT value); probabilities enter as log-probabilities."""
    from oracle.dv3_oracle import twohot_mean
    from oracle.ops_emul import EmulOps

    cases = [({7: 0.7, 8: 0.3}, 2.3), ({8: 0.6, 9: 0.4}, 3.4), ({7: 1.0}, 2.0), ({10: 1.0}, 5.0), ({0: 1.0}, -5.0)]
    logits = torch.full((len(cases), 11), -80.0)
    for r, (probs, _) in enumerate(cases):
        for i, p in probs.items():
            logits[r, i] = math.log(p)
    want = torch.tensor([symexp(v) for _, v in cases])
    out = torch.zeros(len(cases))
    EmulOps().twohot_mean(logits, -5.0, 5.0, out)
    assert torch.allclose(out, want, rtol=1e-5)
    assert torch.allclose(twohot_mean(logits, -5.0, 5.0).reshape(-1), want, rtol=1e-5)
