"""Names the reference keeps in `sheeprl/algos/p2e_dv3/utils.py` (`cli.py:151-181` reads them from the algorithm's
package)."""
from sheeprl_b200.algos.dreamer_v3.utils import Moments, prepare_obs  # noqa: F401

AGGREGATOR_KEYS = {
    "Rewards/rew_avg", "Game/ep_len_avg", "Loss/world_model_loss", "Loss/value_loss_task", "Loss/policy_loss_task",
    "Loss/value_loss_exploration", "Loss/policy_loss_exploration", "Loss/observation_loss", "Loss/reward_loss",
    "Loss/state_loss", "Loss/continue_loss", "Loss/ensemble_loss", "State/kl", "State/post_entropy", "State/prior_entropy",
    "Grads/world_model", "Grads/actor_task", "Grads/critic_task", "Grads/actor_exploration", "Grads/critic_exploration",
    "Grads/ensemble", "Rewards/intrinsic", "Values_exploration/predicted_values", "Values_exploration/lambda_values",
}
MODELS_TO_REGISTER = {"world_model", "ensembles", "actor_exploration", "critic_exploration", "target_critic_exploration",
                      "moments_exploration", "actor_task", "critic_task", "target_critic_task", "moments_task"}
