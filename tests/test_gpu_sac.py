"""SAC on the GPU through the C-ABI: every new kernel against the torch test double, the whole update against the
executed reference (tests/golden/sac_*.pt), and the CUDA-graph capture of the update against eager launches.
Tolerance: 1e-4 relative fp32 (north star); parameters follow tests/test_sac_cpu.assert_group_close."""
import pytest
import torch

from oracle.ops_emul import EmulOps
from tests.test_sac_cpu import check_engine, load, make_engine

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cu():
    from sheeprl_b200.lib import CudaOps

    return CudaOps()


def rnd(*shape, seed=0, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale


def close(a, b, rtol=1e-4, atol=1e-5, what=""):
    a, b = a.detach().cpu(), b.detach().cpu()
    err = (a - b).abs()
    assert bool((err <= atol + rtol * b.abs()).all()), (what, float(err.max()))


@pytest.mark.parametrize("nets,M,N,K", [(1, 256, 256, 17), (2, 256, 256, 23), (2, 256, 1, 256), (2, 8, 16, 8),
                                        (1, 77, 45, 130), (2, 2048, 64, 300), (1, 25600, 32, 768)])
@pytest.mark.parametrize("epi", ["none", "relu", "tanh", "drelu", "dtanh"])
def test_bgemm_forward_layouts(cu, nets, M, N, K, epi):
    em = EmulOps()
    A = rnd(1 if nets > 1 and M % 2 == 0 else nets, M, K, seed=1)            # shared input across nets
    W = rnd(nets, N, K, seed=2, scale=K ** -0.5)
    bias = rnd(nets, N, seed=3)
    aux = torch.tanh(rnd(nets, M, N, seed=4))
    want = torch.zeros(nets, M, N)
    em.bgemm(A, W.transpose(1, 2), want, bias=bias, aux=aux, epi=epi)
    got = torch.zeros(nets, M, N, device="cuda")
    cu.bgemm(A.cuda(), W.cuda().transpose(1, 2), got, bias=bias.cuda(), aux=aux.cuda(), epi=epi)
    close(got, want, what=f"bgemm NT {epi}")


def test_bgemm_backward_layouts_and_rsum(cu):
    em = EmulOps()
    nets, B, H, I = 2, 256, 64, 23
    dY, X, W = rnd(nets, B, H, seed=1), rnd(1, B, I, seed=2), rnd(nets, H, I, seed=3)
    # weight gradient (TN) + bias gradient, accumulate on top of existing values
    dW0, db0 = rnd(nets, H, I, seed=4), rnd(nets, H, seed=5)
    want_w, want_b = dW0.clone(), db0.clone()
    em.bgemm(dY.transpose(1, 2), X, want_w, rsum=want_b, accumulate=True)
    got_w, got_b = dW0.cuda(), db0.cuda()
    cu.bgemm(dY.cuda().transpose(1, 2), X.cuda(), got_w, rsum=got_b, accumulate=True)
    close(got_w, want_w, what="dW"), close(got_b, want_b, what="db")
    # input gradient (NN) restricted to a column range of W (strided B operand)
    want = torch.zeros(nets, B, I - 17)
    em.bgemm(dY, W[:, :, 17:], want)
    got = torch.zeros(nets, B, I - 17, device="cuda")
    cu.bgemm(dY.cuda(), W.cuda()[:, :, 17:], got)
    close(got, want, what="dX cols")


def test_sac_elementwise_kernels(cu):
    em = EmulOps()
    B, A, n = 256, 6, 2
    head = rnd(B, 2 * A, seed=1) * 2.5                       # some log_std outside [-5, 2]: clamp branches
    eps, scale, abias = rnd(B, A, seed=2), torch.rand(A) + 0.5, rnd(A, seed=3)
    la = torch.tensor([-0.3])
    act_w, lp_w, th_w = torch.zeros(B, 9)[:, 3:], torch.zeros(B), torch.zeros(B, A)
    em.sac_sample_fwd(head, eps, scale, abias, act_w, lp_w, th_w)
    xbuf = torch.zeros(B, 9, device="cuda")
    lp_g, th_g = torch.zeros(B, device="cuda"), torch.zeros(B, A, device="cuda")
    cu.sac_sample_fwd(head.cuda(), eps.cuda(), scale.cuda(), abias.cuda(), xbuf[:, 3:], lp_g, th_g)
    close(xbuf[:, 3:], act_w, what="action"), close(lp_g, lp_w, what="logp"), close(th_g, th_w, what="tanh")
    assert float(xbuf[:, :3].abs().max()) == 0.0
    dact = rnd(n, B, A, seed=5)
    dh_w, dh_g = torch.zeros(B, 2 * A), torch.zeros(B, 2 * A, device="cuda")
    em.sac_sample_bwd(head, eps, th_w, scale, dact, la, dh_w)
    cu.sac_sample_bwd(head.cuda(), eps.cuda(), th_g, scale.cuda(), dact.cuda(), la.cuda(), dh_g)
    close(dh_g, dh_w, what="dhead")
    q, r, d = rnd(n, B, seed=6), rnd(B, seed=7), (torch.rand(B) < 0.1).float()
    y_w, y_g = torch.zeros(B), torch.zeros(B, device="cuda")
    em.sac_target(q, lp_w, r, d, la, 0.99, y_w)
    cu.sac_target(q.cuda(), lp_g, r.cuda(), d.cuda(), la.cuda(), 0.99, y_g)
    close(y_g, y_w, what="target")
    dq_w, l_w, dq_g, l_g = torch.zeros(n, B), torch.zeros(1), torch.zeros(n, B, device="cuda"), torch.zeros(1, device="cuda")
    em.sac_critic_loss(q, y_w, dq_w, l_w)
    cu.sac_critic_loss(q.cuda(), y_g, dq_g, l_g)
    close(dq_g, dq_w, what="dq"), close(l_g, l_w, what="qf loss")
    outs_w = [torch.zeros(n, B), torch.zeros(1), torch.zeros(1), torch.zeros(1)]
    outs_g = [t.cuda() for t in outs_w]
    em.sac_actor_loss(q, lp_w, la, -6.0, *outs_w)
    cu.sac_actor_loss(q.cuda(), lp_g, la.cuda(), -6.0, *outs_g)
    for g, w, nme in zip(outs_g, outs_w, ("dq_pi", "actor loss", "alpha loss", "dlog_alpha")):
        close(g, w, what=nme)


def test_fill_normal_statistics(cu):
    a, b = torch.empty(1 << 20, device="cuda"), torch.empty(1 << 20, device="cuda")
    ctr = torch.zeros(1, dtype=torch.int32, device="cuda")
    cu.fill_normal(a, 7, 1, ctr)
    cu.fill_normal(b, 7, 2, ctr)
    assert abs(float(a.mean())) < 5e-3 and abs(float(a.var()) - 1.0) < 1e-2 and abs(float((a ** 4).mean()) - 3.0) < 0.1
    assert abs(float((a * b).mean())) < 5e-3 and not torch.equal(a, b)
    c = torch.empty_like(a)
    cu.fill_normal(c, 7, 1, ctr)
    assert torch.equal(a, c)                                  # counter-keyed: same key, same draw
    ctr += 1
    cu.fill_normal(c, 7, 1, ctr)
    assert not torch.equal(a, c)


@pytest.mark.parametrize("name", ["sac_tiny", "sac_c4"])
def test_engine_matches_reference(cu, name):
    fx = load(name)
    check_engine(fx, make_engine(fx, device="cuda", ops=cu), name)


def test_update_is_graph_capturable(cu):
    """whole update (no host sync inside) replayed from a CUDA graph == eager launches"""
    fx = load("sac_c4")
    e1, e2 = make_engine(fx, "cuda", cu), make_engine(fx, "cuda", cu)
    st = fx["steps"][0]
    data = {k: v.cuda() for k, v in st["data"].items()}
    noise = {"eps_next": st["eps_next"].cuda(), "eps_cur": st["eps_cur"].cuda()}
    for _ in range(3):
        e1.train_step(data, True, noise)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        e2.train_step(data, True, noise)                     # warm-up on the side stream (step 1)
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        e2.train_step(data, True, noise)                     # captured: step 2 happens at first replay
    g.replay()
    g.replay()
    torch.cuda.synchronize()
    a, b = e1.export_reference_state(), e2.export_reference_state()
    for grp in a:
        for k in a[grp]:
            close(b[grp][k], a[grp][k], rtol=1e-5, atol=1e-6, what=f"graph {grp}/{k}")
