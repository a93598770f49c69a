"""B200 twin of `colosseum/experiment/indicators.py:9-45`: the episodic regret indicators `MDPLoop` evaluates at every
log tick (experiment/agent_mdp_interaction.py:518-578) -- one episodic policy evaluation (and, the first time, one
backward induction) on the GPU."""
import numpy as np

from .dynamic_programming import _is_tensor, episodic_policy_evaluation, episodic_value_iteration


def _host(x):
    return x.cpu().numpy() if _is_tensor(x) else np.asarray(x)


def get_episodic_regret_at_time_zero(H, T, R, policy, optimal_value=None):
    """indicators.py:9-26: optimal_value[0] - V_policy[0]."""
    assert T.ndim == 3 if not _is_tensor(T) else T.dim() == 3, "We don't need the episodic transition matrix here."
    _, V = episodic_policy_evaluation(H, T, R, policy)
    if optimal_value is None:
        _, optimal_value = episodic_value_iteration(H, T, R)
    return optimal_value[0] - V[0]


def get_episodic_regrets_and_average_reward_at_time_zero(H, T, R, policy, starting_state_distribution,
                                                         optimal_value=None):
    """indicators.py:29-45: (max(optimal_value[0] - V[0], 0), sum(V[0] * starting_state_distribution))."""
    _, V = episodic_policy_evaluation(H, T, R, policy)
    V0 = _host(V[0])
    avg = float((V0 * _host(starting_state_distribution)).sum())
    if optimal_value is None:
        _, optimal_value = episodic_value_iteration(H, T, R)
    return np.maximum(_host(optimal_value[0]) - V0, 0.0), avg
