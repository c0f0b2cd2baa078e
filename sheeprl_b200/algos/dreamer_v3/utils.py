"""Dreamer-V3 side objects the reference keeps in `sheeprl/algos/dreamer_v3/utils.py`.

`Moments` here is a thin state holder: the percentile / EMA arithmetic runs in the
`b200rl_moments_update` kernel inside `DV3Engine` (reference: dreamer_v3/utils.py:40-63)."""
from __future__ import annotations

import torch

AGGREGATOR_KEYS = {
    "Rewards/rew_avg", "Game/ep_len_avg", "Loss/world_model_loss", "Loss/value_loss", "Loss/policy_loss",
    "Loss/observation_loss", "Loss/reward_loss", "Loss/state_loss", "Loss/continue_loss", "State/kl",
    "State/post_entropy", "State/prior_entropy", "Grads/world_model", "Grads/actor", "Grads/critic",
}
MODELS_TO_REGISTER = {"world_model", "actor", "critic", "target_critic", "moments"}


class Moments(torch.nn.Module):
    """Buffers `low` / `high` (checkpoint-compatible with the reference's Moments.state_dict())."""

    def __init__(self, decay: float = 0.99, max_: float = 1e8, percentile_low: float = 0.05,
                 percentile_high: float = 0.95) -> None:
        super().__init__()
        self._decay, self._max = decay, max_
        self._percentile_low, self._percentile_high = percentile_low, percentile_high
        self.register_buffer("low", torch.zeros((), dtype=torch.float32))
        self.register_buffer("high", torch.zeros((), dtype=torch.float32))

    def bind(self, state: torch.Tensor) -> None:
        """Alias the buffers onto the engine's 2-float device state so the kernel updates them in place."""
        with torch.no_grad():
            state[0] = self.low.to(state.device)
            state[1] = self.high.to(state.device)
        self.low = state[0]
        self.high = state[1]


def prepare_obs(fabric, obs, *, cnn_keys=(), num_envs: int = 1, **kwargs):
    """numpy observations -> device tensors `[1, num_envs, ...]` (reference: dreamer_v3/utils.py:80-91).  Image keys stay
    uint8: `PlayerDV3.get_actions` normalises them in the obs_prep kernel, so the host->device copy is 4x smaller than
    the reference's float32 `/255 - 0.5` tensor."""
    out = {}
    for k, v in obs.items():
        t = torch.from_numpy(v.copy()).to(fabric.device)
        if k in cnn_keys:
            t = t.view(1, num_envs, -1, *v.shape[-2:])
            out[k] = t if t.dtype == torch.uint8 else t.float() / 255 - 0.5
        else:
            out[k] = t.float().view(1, num_envs, -1)
    return out
