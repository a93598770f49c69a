"""Pins the oracle's Boltzmann exploration (orc_boltzmann_action) to the REFERENCE's QValuesActor.

Run in the build container (needs /root/reference):  python tests/golden/make_actor_golden.py
The unmodified colosseum.agent.actors.Q_values_actor.QValuesActor is built with a temperature SCHEDULE (a function of its
interaction counter: the only form the reference can run, its float branch wraps the number into a lambda that returns
itself, Q_values_actor.py:44-49) and asked for actions on random q-tables.  Before every call the uniform its
`self._rng.choice(..., p=...)` is about to consume is read off a copy of the RandomState (choice with p draws one
random_sample and searches the normalised cumulative sum).  tests/golden/actor.npz keeps (Q, state, temperature, u,
action); tests/test_oracle_agents.py feeds (q row, temperature, u) to the oracle and expects the same actions.
"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_qlearning_golden import reference_models  # noqa: E402
from oracle import oracle as orc  # noqa: E402

CASES = [("a2", 12, 2, lambda t: 0.5 + 0.0005 * t, 3.0), ("a4", 9, 4, lambda t: 2.0 / (1 + 0.001 * t), 1.0),
         ("a6_sharp", 7, 6, lambda t: 5.0 + 0.002 * t, 0.5)]
N_CALLS, SEED = 4000, 3


def main():
    reference_models()
    import importlib

    qa = importlib.import_module("colosseum.agent.actors.Q_values_actor")
    out = {}
    for name, S, A, temp, scale in CASES:
        rng = np.random.RandomState(SEED)
        Q = (rng.randn(S, A) * scale).astype(np.float32)
        spec = types.SimpleNamespace(observations=types.SimpleNamespace(num_values=S),
                                     actions=types.SimpleNamespace(num_values=A))
        actor = qa.QValuesActor(SEED, spec, None, temp)
        actor.set_q_values(Q)
        states = rng.randint(S, size=N_CALLS)
        us, acts, temps = [], [], []
        for k, s in enumerate(states):
            peek = np.random.RandomState()
            peek.set_state(actor._rng.get_state())
            us.append(peek.random_sample())
            temps.append(float(temp(k + 1)))  # _total_interactions is incremented before it is used (:61)
            acts.append(int(actor.select_action(types.SimpleNamespace(observation=int(s)), k)))
        out[f"{name}.Q"], out[f"{name}.states"] = Q, states.astype(np.int32)
        out[f"{name}.u"], out[f"{name}.temp"], out[f"{name}.action"] = np.asarray(us), np.asarray(temps), np.asarray(acts, np.int32)
        ours = np.array([orc.boltzmann_action(Q[s], t, u) for s, t, u in zip(states, temps, us)])
        print(f"{name}: {N_CALLS} calls, oracle == reference actor on {(ours == out[f'{name}.action']).mean():.6f}, "
              f"action histogram {np.bincount(acts, minlength=A).tolist()}")
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "actor.npz"), **out)
    print("wrote tests/golden/actor.npz")


if __name__ == "__main__":
    main()
