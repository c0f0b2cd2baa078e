"""TEST INFRASTRUCTURE (oracle) — CPU restatements of the PPO pieces on the hot path.

`gae_oracle` restates `sheeprl/utils/utils.py:63-100` (`gae`).  Pinned against the imported reference in
tests/test_oracle_pin.py (container only) and against tests/golden/gae_small.pt everywhere.
"""
from __future__ import annotations

import torch


def gae_oracle(rewards, values, dones, next_value, num_steps, gamma, gae_lambda):
    """Generalised advantage estimation, reverse scan over `num_steps` (reference: utils/utils.py:63-100).
    `dones[t]` masks the bootstrap from step t+1 into step t; the last step bootstraps `next_value`."""
    alive = (dones == 0).to(rewards.dtype)
    adv = torch.zeros_like(rewards)
    running = torch.zeros_like(rewards[0])
    for t in range(num_steps - 1, -1, -1):
        if t == num_steps - 1:
            nv, mask = next_value, alive[-1]
        else:
            nv, mask = values[t + 1], alive[t]
        delta = rewards[t] + nv * mask * gamma - values[t]
        running = delta + mask * running * gamma * gae_lambda
        adv[t] = running
    return adv + values, adv
