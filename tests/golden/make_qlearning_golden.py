"""Pins the Q-learning update rules of the oracle (oracle.QLearningLoops) to the REFERENCE's model classes.

Run in the build container (needs /root/reference):  python tests/golden/make_qlearning_golden.py
For each case the oracle runs N loops with a trace; every loop's (s_t, a_t, obs_tp1, r, h) sequence is then replayed
through the unmodified reference classes
    colosseum.agent.agents.episodic.q_learning.QValuesModel          (Hoeffding and Bernstein)
    colosseum.agent.agents.infinite_horizon.q_learning._QValuesModel
and the reference's final tables are stored in tests/golden/qlearning.npz next to the trace.  The reference modules are
imported without their package __init__ (which pulls sonnet/tensorflow) and with a stub for ray.tune (only used by
hyper-parameter search helpers).
"""
import importlib
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_instance  # noqa: E402
from colosseum_b200.tables import MDPTables  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from oracle.reference_import import REFERENCE_ROOT, import_reference  # noqa: E402

CASES = [  # name, instance, kwargs of oracle.QLearningLoops
    ("epi_hoeffding", "c1_riverswim_epi", dict(optimization_horizon=2000, p=0.05, c_1=0.7, min_at=0.0, UCB_type="hoeffding")),
    ("epi_bernstein", "frozenlake4_epi", dict(optimization_horizon=3000, p=0.05, c_1=0.4, c_2=0.9, min_at=0.05, UCB_type="bernstein")),
    ("epi_bernstein_taxi", "taxi_epi", dict(optimization_horizon=3000, p=0.05, c_1=1.0, c_2=0.3, min_at=0.0, UCB_type="bernstein", epsilon_greedy=0.1)),
    # min_at large enough to win python's max() from the first visits on: alpha_t is then a PYTHON float and the
    # update runs in float32 (NEP 50) -- q_learning.py:66, :92-102
    ("epi_hoeffding_minat", "c1_riverswim_epi", dict(optimization_horizon=2000, p=0.05, c_1=0.7, min_at=0.5, UCB_type="hoeffding")),
    ("epi_bernstein_minat", "frozenlake4_epi", dict(optimization_horizon=3000, p=0.05, c_1=0.4, c_2=0.9, min_at=0.4, UCB_type="bernstein")),
    ("cont", "frozenlakecontinuous_ergo0", dict(optimization_horizon=5000, min_at=0.02, confidence=0.95, span_approx_weight=0.6, h_weight=0.8)),
    ("cont_taxi", "taxicontinuous_ergo0", dict(optimization_horizon=4000, min_at=0.0, confidence=0.9, span_approx_weight=1.0, h_weight=1.0, epsilon_greedy=0.05)),
]
N_LOOPS, N_STEPS, SEED = 3, 600, 17


def host_tables(tb):
    return orc.HostTables(tb.S, tb.A, H=tb.H, succ_cum=tb.succ_cum, succ_idx=tb.succ_idx, succ_len=tb.succ_len,
                          rew_cls_succ=tb.rew_cls_succ, rew_q=tb.rew_q, rmin=tb.rmin, rmax=tb.rmax,
                          start_cum=tb.start_cum, start_idx=tb.start_idx)


def reference_models():
    import_reference()
    ray, tune = types.ModuleType("ray"), types.ModuleType("ray.tune")
    search, sample = types.ModuleType("ray.tune.search"), types.ModuleType("ray.tune.search.sample")
    sample.Domain = type("Domain", (), {})
    search.sample, tune.search, ray.tune = sample, search, tune
    tune.uniform = lambda a, b: (a, b)
    sys.modules.update({"ray": ray, "ray.tune": tune, "ray.tune.search": search, "ray.tune.search.sample": sample})
    for n in ("colosseum.agent", "colosseum.agent.agents", "colosseum.agent.agents.episodic",
              "colosseum.agent.agents.infinite_horizon"):
        if n not in sys.modules:
            m = types.ModuleType(n)
            m.__path__ = [os.path.join(REFERENCE_ROOT, *n.split("."))]
            sys.modules[n] = m
    qe = importlib.import_module("colosseum.agent.agents.episodic.q_learning")
    qc = importlib.import_module("colosseum.agent.agents.infinite_horizon.q_learning")
    return qe.QValuesModel, qc._QValuesModel, qc.get_H


def main():
    QE, QC, ref_get_H = reference_models()
    out = {}
    for name, inst, kw in CASES:
        tb = MDPTables.from_golden(load_instance(inst))
        loops = orc.QLearningLoops(host_tables(tb), N_LOOPS, seed=SEED, **kw)
        start = loops.state.copy()
        trace = loops.steps(N_STEPS, trace=True)
        spec = types.SimpleNamespace(observations=types.SimpleNamespace(num_values=tb.S),
                                     actions=types.SimpleNamespace(num_values=tb.A), rewards_range=(tb.rmin, tb.rmax),
                                     time_horizon=tb.H if tb.H > 0 else np.inf)
        tabs = {}
        for i in range(N_LOOPS):
            if tb.H > 0:
                m = QE(SEED, spec, kw["optimization_horizon"], kw["p"], kw["c_1"], kw.get("c_2"), kw["min_at"], kw["UCB_type"])
            else:
                m = QC(SEED, spec, kw["optimization_horizon"], kw["min_at"], kw["confidence"], kw["span_approx_weight"],
                       None, kw["h_weight"], ref_get_H)
                assert abs(m.H - loops.H_eff) <= 1e-12 * m.H, (m.H, loops.H_eff)
            h = 0
            for k in range(N_STEPS):
                s, a, obs, rbits = (int(x) for x in trace[k, i])
                r = float(np.int32(rbits).view(np.float32))
                m.step_update(types.SimpleNamespace(observation=s), a,
                              types.SimpleNamespace(observation=obs, reward=r), h)
                h = 0 if obs < 0 else h + 1
            for f in ("N", "Q", "V", "Q_main", "mu", "sigma", "beta"):
                if hasattr(m, f):
                    tabs.setdefault(f, []).append(np.asarray(getattr(m, f)))
        out[f"{name}.start"] = start
        out[f"{name}.trace"] = trace
        for f, v in tabs.items():
            out[f"{name}.ref_{f}"] = np.stack(v)
        ours = {"N": loops.cnt, "Q": loops.Q, "V": loops.V}
        for f in ("Q_main", "mu", "sigma", "beta"):
            if hasattr(loops, f):
                ours[f] = getattr(loops, f)
        for f, v in ours.items():
            ref = out[f"{name}.ref_{f}"]
            err = np.abs(v.astype(np.float64) - ref.astype(np.float64)).max() / max(1.0, np.abs(ref).max())
            print(f"{name:20s} {f:7s} max rel err vs reference {err:.2e}  exact={np.array_equal(v, ref)}")
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "qlearning.npz"), **out)
    print("wrote tests/golden/qlearning.npz")


if __name__ == "__main__":
    main()
