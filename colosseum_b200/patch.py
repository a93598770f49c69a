"""Install the B200 path behind the reference's own names (drop-in wiring).

    import colosseum                      # the unmodified reference
    import colosseum_b200.patch as patch
    patch.install()                       # colosseum.dynamic_programming.* / colosseum.hardness.measures.* -> GPU
    mdp.diameter, mdp.value_norm, mdp.optimal_value_functions ...   # reference code, GPU arithmetic
    patch.uninstall()

The reference binds these functions by name at import time (`from colosseum.dynamic_programming import
discounted_value_iteration` in mdp/base.py:15-17, mdp/base_finite.py:7-11, hardness/measures/diameter.py:12-13,
hardness/measures/value_norm.py:7-9, experiment/indicators.py:5-6, agent/agents/*/posterior_sampling.py:13,
agent/agents/infinite_horizon/ucrl2.py:12-13), so `install()` rebinds every `colosseum.*` module attribute that
IS one of the original function objects.  Nothing is dispatched at run time: after install() the GPU functions are
the only implementation those names refer to.
"""
import sys

from . import dynamic_programming as _dp
from . import hardness as _hd
from . import indicators as _ind
from . import markov_chain as _mc

# (module that defines the original, attribute name) -> replacement
REPLACEMENTS = {
    ("colosseum.dynamic_programming.finite_horizon", "episodic_value_iteration"): _dp.episodic_value_iteration,
    ("colosseum.dynamic_programming.finite_horizon", "episodic_policy_evaluation"): _dp.episodic_policy_evaluation,
    ("colosseum.dynamic_programming.infinite_horizon", "discounted_value_iteration"): _dp.discounted_value_iteration,
    ("colosseum.dynamic_programming.infinite_horizon", "discounted_policy_evaluation"): _dp.discounted_policy_evaluation,
    ("colosseum.dynamic_programming.infinite_horizon", "discounted_policy_iteration"): _dp.discounted_policy_iteration,
    ("colosseum.dynamic_programming.infinite_horizon", "extended_value_iteration"): _dp.extended_value_iteration,
    ("colosseum.hardness.measures.diameter", "get_diameter"): _hd.get_diameter,
    ("colosseum.hardness.measures.value_norm", "calculate_norm_discounted"): _hd.calculate_norm_discounted,
    ("colosseum.hardness.measures.value_norm", "calculate_norm_average"): _hd.calculate_norm_average,
    ("colosseum.hardness.measures.sum_reciprocals_suboptimality_gaps", "get_sum_reciprocals_suboptimality_gaps"):
        _hd.get_sum_reciprocals_suboptimality_gaps,
    # average-reward helpers (mdp/utils/markov_chain.py:12-137) and the episodic regret indicators
    # (experiment/indicators.py:9-45; that package needs ray & co. to import: skipped when it cannot be imported)
    ("colosseum.mdp.utils.markov_chain", "get_average_reward"): _mc.get_average_reward,
    ("colosseum.mdp.utils.markov_chain", "get_average_rewards"): _mc.get_average_rewards,
    ("colosseum.mdp.utils.markov_chain", "get_transition_probabilities"): _mc.get_transition_probabilities,
    ("colosseum.mdp.utils.markov_chain", "get_stationary_distribution"): _mc.get_stationary_distribution,
    ("colosseum.experiment.indicators", "get_episodic_regret_at_time_zero"): _ind.get_episodic_regret_at_time_zero,
    ("colosseum.experiment.indicators", "get_episodic_regrets_and_average_reward_at_time_zero"):
        _ind.get_episodic_regrets_and_average_reward_at_time_zero,
}

_saved = []  # (module, attribute, original object)


def install(prefix="colosseum", reference_iterates=True):
    """Rebind the reference's hot-path entry points to the GPU implementations.  Returns the number of bindings
    replaced.  Needs the reference package to be imported already (its modules are found in sys.modules).
    reference_iterates=True (the default: a drop-in returns the reference's numbers) makes the discounted solvers
    and the continuous diameter sweep in place like the reference's numba kernels
    (`dynamic_programming.set_sweep_order("gauss_seidel")`): they stop where the reference stops and return its own
    early-stopped iterates -- at the reference's default epsilon = 1e-3 a synchronous iterate differs from them by up
    to ~0.05 absolute, enough to flip greedy policies on near-ties (ucrl2.py:82, posterior_sampling.py:177/:374).
    reference_iterates=False selects synchronous sweeps iterated to the fixed point."""
    _dp.set_sweep_order("gauss_seidel" if reference_iterates else "jacobi")
    if _saved:
        return len(_saved)
    originals = {}
    for (mod_name, attr), repl in REPLACEMENTS.items():
        mod = sys.modules.get(mod_name)
        if mod is None:
            try:
                __import__(mod_name)
            except Exception:  # optional parts of the reference whose own dependencies are absent
                continue
            mod = sys.modules[mod_name]
        originals[id(getattr(mod, attr))] = repl
    for name, mod in list(sys.modules.items()):
        if mod is None or not (name == prefix or name.startswith(prefix + ".")):
            continue
        for attr, obj in list(vars(mod).items()):
            repl = originals.get(id(obj))
            if repl is not None and callable(obj):
                _saved.append((mod, attr, obj))
                setattr(mod, attr, repl)
    return len(_saved)


def uninstall():
    """Restore every binding `install()` replaced."""
    n = len(_saved)
    _dp.set_sweep_order("jacobi")
    while _saved:
        mod, attr, obj = _saved.pop()
        setattr(mod, attr, obj)
    return n
