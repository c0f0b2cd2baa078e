"""Plan2Explore (Dreamer-V3) exploration update on the GPU through the C-ABI: the engine against the executed
reference (tests/golden/p2e_tiny.pt), and the public build_agent()/train() surface on the same fixture."""
import pytest
import torch

from tests.test_p2e_cpu import check_engine, check_public_api, load, make_engine

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cu():
    from sheeprl_b200.lib import CudaOps

    return CudaOps()


def test_engine_matches_reference(cu):
    fx, cfg = load()
    check_engine(fx, cfg, make_engine(fx, cfg, device="cuda", ops=cu))


def test_public_api_matches_reference(cu):
    check_public_api(device="cuda", ops=cu)


def test_production_noise_step_is_finite_and_explores(cu):
    """on-device Philox noise: two updates stay finite, the ensemble members disagree (intrinsic reward > 0) and the
    exploration and task rollouts use different noise streams"""
    fx, cfg = load()
    eng = make_engine(fx, cfg, device="cuda", ops=cu)
    data = {k: v.clone().float().cuda() for k, v in fx["data"][0].items()}
    for _ in range(2):
        eng.train_step(data, None)
    md = {k: float(v) for k, v in eng.metrics_dict().items()}
    assert all(v == v and abs(v) < 1e30 for v in md.values()), md
    assert md["Rewards/intrinsic_intrinsic"] > 0
    assert not torch.equal(eng.noise_img_state_expl, eng.noise_img_state)
