"""PPO side constants the CLI looks up in `<algo module>.utils` (sheeprl/cli.py:151-181; reference:
sheeprl/algos/ppo/utils.py:18-20)."""
AGGREGATOR_KEYS = {"Rewards/rew_avg", "Game/ep_len_avg", "Loss/value_loss", "Loss/policy_loss", "Loss/entropy_loss"}
MODELS_TO_REGISTER = {"agent"}
