"""Pins the posterior updates of the oracle's PSRL loops (oracle.PSRLLoops) to the REFERENCE's conjugate models.

Run in the build container (needs /root/reference):  python tests/golden/make_psrl_golden.py
The oracle runs N loops acting epsilon-greedily on the optimal Q of the true MDP; every loop's transitions are replayed
through the unmodified `BayesianMDPModel.step_update` (N_NIG + M_DIR, the PSRL defaults); the reference's posterior
parameters are stored in tests/golden/psrl.npz next to the trace.
"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from conftest import load_instance  # noqa: E402
from colosseum_b200.tables import MDPTables  # noqa: E402
from make_qlearning_golden import host_tables  # noqa: E402
from oracle import oracle as orc  # noqa: E402

CASES = [("c1_riverswim_epi", "c1_riverswim_epi", dict(epsilon_greedy=0.3)),
         ("frozenlake4_epi", "frozenlake4_epi", dict(epsilon_greedy=0.3, rewards_prior_prms=[0.5, 2, 1.5, 3],
                                                      transitions_prior_prms=[0.25])),
         ("taxi_epi", "taxi_epi", dict(epsilon_greedy=0.5)),
         ("frozenlake4_nn", "frozenlake4_epi", dict(epsilon_greedy=0.3, reward_prior_model="N_N",
                                                     rewards_prior_prms=[0.3, 2.0])),
         ("c1_riverswim_nn", "c1_riverswim_epi", dict(epsilon_greedy=0.4, reward_prior_model="N_N",
                                                      rewards_prior_prms=[1.0, 1.0]))]
N_LOOPS, N_EPISODES, SEED = 3, 40, 5


def optimal_q(g, tb):
    Q, _ = orc.episodic_f32(tb.H, np.asarray(g["T"], np.float32), np.asarray(g["R"], np.float32))
    return np.asarray(Q, np.float32)


def main():
    from oracle.reference_import import import_reference

    import_reference()
    from colosseum.agent.mdp_models.bayesian_model import BayesianMDPModel
    from colosseum.agent.mdp_models.bayesian_models import RewardsConjugateModel, TransitionsConjugateModel

    out = {}
    for name, inst, kw in CASES:
        g = load_instance(inst)
        tb = MDPTables.from_golden(g)
        loops = orc.PSRLLoops(host_tables(tb), N_LOOPS, seed=SEED, **kw)
        loops.set_q(optimal_q(g, tb))
        trace = loops.steps(N_EPISODES * tb.H, trace=True)
        spec = types.SimpleNamespace(observations=types.SimpleNamespace(num_values=tb.S),
                                     actions=types.SimpleNamespace(num_values=tb.A), rewards_range=(tb.rmin, tb.rmax),
                                     time_horizon=tb.H)
        nig, dirs = [], []
        for i in range(N_LOOPS):
            rmodel = RewardsConjugateModel.N_N if kw.get("reward_prior_model") == "N_N" else RewardsConjugateModel.N_NIG
            m = BayesianMDPModel(SEED, spec,
                                 reward_prior_model=rmodel if "rewards_prior_prms" in kw else None,
                                 transitions_prior_model=TransitionsConjugateModel.M_DIR if "transitions_prior_prms" in kw else None,
                                 rewards_prior_prms=kw.get("rewards_prior_prms"),
                                 transitions_prior_prms=kw.get("transitions_prior_prms"))
            for k in range(trace.shape[0]):
                s, a, obs, rbits = (int(x) for x in trace[k, i])
                r = float(np.int32(rbits).view(np.float32))
                last = obs < 0
                m.step_update(types.SimpleNamespace(observation=s), a,
                              types.SimpleNamespace(observation=obs, reward=r, last=lambda last=last: last), 0)
            nig.append(np.asarray(m._rewards_model.hyper_params))
            dirs.append(np.asarray(m._transitions_model.hyper_params))
        out[f"{name}.trace"] = trace
        out[f"{name}.ref_nig"] = np.stack(nig)  # [N,S,A,4] for N_NIG, [N,S,A,2] for N_N
        out[f"{name}.ref_dir"] = np.stack(dirs)
        k = out[f"{name}.ref_nig"].shape[-1]
        for what, ours, ref in (("rew", loops.nig_hyper[..., :k], out[f"{name}.ref_nig"]),
                                ("dir", loops.dir_hyper, out[f"{name}.ref_dir"])):
            err = np.abs(ours.astype(np.float64) - ref).max() / max(1.0, np.abs(ref).max())
            print(f"{name:20s} {what}: max rel err vs reference {err:.2e} exact={np.array_equal(ours, ref.astype(ours.dtype))} dtype {ref.dtype}")
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "psrl.npz"), **out)
    print("wrote tests/golden/psrl.npz")


if __name__ == "__main__":
    main()
