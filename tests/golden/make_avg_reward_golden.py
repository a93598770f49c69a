"""Generate tests/golden/avg_reward.npz by running the UNMODIFIED reference (build container only): for a handful of
continuous MDP instances, the optimal / worst / random policies, the Markov chains they induce, their stationary
distributions and average rewards (colosseum/mdp/base.py:767-941 -> colosseum/mdp/utils/markov_chain.py:12-137)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.reference_import import import_reference  # noqa: E402

import_reference()
import colosseum.mdp  # noqa: E402,F401
from colosseum.mdp.deep_sea import DeepSeaContinuous  # noqa: E402
from colosseum.mdp.frozen_lake import FrozenLakeContinuous  # noqa: E402
from colosseum.mdp.minigrid_empty import MiniGridEmptyContinuous  # noqa: E402
from colosseum.mdp.river_swim import RiverSwimContinuous  # noqa: E402
from colosseum.mdp.simple_grid import SimpleGridContinuous  # noqa: E402
from colosseum.mdp.taxi import TaxiContinuous  # noqa: E402
from colosseum.mdp.utils.markov_chain import get_average_rewards, get_transition_probabilities  # noqa: E402


def main():
    cases = {
        "doc_simplegrid4": SimpleGridContinuous(seed=0, size=4, p_rand=0.01, n_starting_states=3),
        "deepsea10_prand": DeepSeaContinuous(seed=0, size=10, p_rand=0.1),
        "riverswim20": RiverSwimContinuous(seed=0, size=20, p_rand=0.05, p_lazy=0.1),
        "frozenlake5": FrozenLakeContinuous(seed=0, size=5, p_frozen=0.9, p_rand=0.1),
        "minigrid6": MiniGridEmptyContinuous(seed=0, size=6, p_rand=0.1, p_lazy=0.05),
        "taxi4": TaxiContinuous(seed=0, size=4, length=1, width=1, space=1, n_locations=3, p_rand=0.1),
    }
    out = {"names": np.array(list(cases))}
    for name, mdp in cases.items():
        T, R = mdp.T, mdp.R
        pols = {"opt": mdp.get_optimal_policy(True), "worst": mdp.get_worst_policy(True), "rand": mdp.random_policy}
        ssp = list(mdp.starting_states_and_probs)
        out[f"{name}_T"], out[f"{name}_R"] = T, R
        out[f"{name}_start_idx"] = np.asarray([s for s, _ in ssp], np.int32)
        out[f"{name}_start_prob"] = np.asarray([p for _, p in ssp], np.float64)
        for k, pi in pols.items():
            sd = mdp.get_stationary_distribution(pi)
            out[f"{name}_{k}_pi"] = np.asarray(pi, np.float32)
            out[f"{name}_{k}_sd"] = np.asarray(sd, np.float64)
            out[f"{name}_{k}_tps"] = get_transition_probabilities(T, pi)
            out[f"{name}_{k}_rs"] = get_average_rewards(R, pi)
        out[f"{name}_opt_ar"] = float(mdp.optimal_average_reward)
        out[f"{name}_worst_ar"] = float(mdp.worst_average_reward)
        out[f"{name}_rand_ar"] = float(mdp.random_average_reward)
        out[f"{name}_undisc_norm"] = float(mdp.undiscounted_value_norm)  # hardness/measures/value_norm.py:64-93
        print(name, T.shape, out[f"{name}_opt_ar"], out[f"{name}_worst_ar"], out[f"{name}_rand_ar"], out[f"{name}_undisc_norm"])
    np.savez_compressed(os.path.join(HERE, "avg_reward.npz"), **out)


if __name__ == "__main__":
    main()
