"""SAC side constants the CLI looks up in `<algo module>.utils` (sheeprl/cli.py:151-181; reference:
sheeprl/algos/sac/utils.py:17-22)."""
AGGREGATOR_KEYS = {"Rewards/rew_avg", "Game/ep_len_avg", "Loss/value_loss", "Loss/policy_loss", "Loss/alpha_loss"}
MODELS_TO_REGISTER = {"agent"}
