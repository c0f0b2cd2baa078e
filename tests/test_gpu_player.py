"""PlayerDV3 through the C-ABI on the GPU, against the executed reference (tests/golden/dv3_player_*.pt); also with raw
uint8 observations (the kernel normalises) and with on-device Philox noise (finite, valid one-hot / bounded actions)."""
import pytest
import torch

from tests.test_player_cpu import check, run_player

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cu():
    from sheeprl_b200.lib import CudaOps

    return CudaOps()


@pytest.mark.parametrize("name", ["dv3_player_discrete", "dv3_player_continuous", "dv3_player_vector", "dv3_player_vector_only"])
@pytest.mark.parametrize("uint8_obs", [False, True])
def test_player_matches_reference(cu, name, uint8_obs):
    fx, got, cont = run_player(name, device="cuda", ops=cu, uint8_obs=uint8_obs)
    check(fx, got, cont)


def test_player_with_device_noise(cu):
    from sheeprl_b200.algos.dreamer_v3.player import PlayerDV3
    from sheeprl_b200.engine import DV3Engine
    from tests.helpers import load_fixture

    tf, cfg = load_fixture("dv3_tiny_a")
    eng = DV3Engine(cfg, tf["actions_dim"], in_channels=3, device="cuda", ops=cu)
    eng.wm.load(tf["init"]["wm"]), eng.actor.load(tf["init"]["actor"])
    player = PlayerDV3(eng, 4)
    player.init_states()
    seen = set()
    for s in range(6):
        obs = torch.randint(0, 256, (1, 4, 3, 64, 64), dtype=torch.uint8, device="cuda")
        acts = player.get_actions({"rgb": obs})
        for a, ad in zip(acts, tf["actions_dim"]):
            assert a.shape == (1, 4, ad) and torch.all(a.sum(-1) == 1)
        z = player.stochastic_state.reshape(4, eng.S, eng.D)
        assert torch.all(z.sum(-1) == 1) and torch.isfinite(player.recurrent_state).all()
        seen.add(tuple(torch.cat(acts, -1).flatten().tolist()))
    assert len(seen) > 1                                        # the Philox counter advances between calls
