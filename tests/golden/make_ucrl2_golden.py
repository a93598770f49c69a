"""Pins the oracle's UCRL2Continuous loops (oracle.UCRL2Loops) to the REFERENCE's agent class.

Run in the build container (needs /root/reference):  python tests/golden/make_ucrl2_golden.py
For each case the oracle runs N complete loops (its own extended value iteration as the planner) with a trace.  Every
loop's (s_t, a_t, s_tp1, r) sequence is then replayed through the UNMODIFIED
    colosseum.agent.agents.infinite_horizon.ucrl2.UCRL2Continuous
-- before_start_interacting, then per step step_update / is_episode_end / episode_end_update, i.e. MDPLoop.run's calls
(experiment/agent_mdp_interaction.py:238-262) with the reference's own numba extended_value_iteration -- and stored in
tests/golden/ucrl2.npz:  the trace, the interaction times at which the reference ended its artificial episodes, its final
model tables (N, P, estimated_rewards, variance_proxy_reward, estimated_holding_times, iteration, episode, delta) and
its Q after the last episode end.  The script prints, and tests/test_oracle_agents.py asserts, that the oracle's episode
ends and model tables are IDENTICAL and its Q within the stopping tolerance of the extended VI.
The reference module is imported without its package __init__ (which pulls sonnet/tensorflow), with a stub for ray.tune.
"""
import importlib
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
from make_qlearning_golden import host_tables, reference_models  # noqa: E402
from oracle import oracle as orc  # noqa: E402

CASES = [  # name, instance, kwargs of oracle.UCRL2Loops / UCRL2Continuous
    ("riverswim_chernoff", "riverswimcontinuous_ergo0", dict(alpha_r=1.0, alpha_p=1.0)),
    ("frozenlake_bernstein", "frozenlakecontinuous_ergo0", dict(alpha_r=0.4, alpha_p=0.7, bound_type_p="bernstein")),
    ("deepsea_eps", "deepsea10", dict(alpha_r=0.2, alpha_p=0.1, epsilon_greedy=0.05)),
]
N_LOOPS, N_STEPS, SEED = 3, 4000, 23


def mdp_spec(tb):
    return types.SimpleNamespace(observations=types.SimpleNamespace(num_values=tb.S),
                                 actions=types.SimpleNamespace(num_values=tb.A),
                                 rewards_range=(float(tb.rmin), float(tb.rmax)), time_horizon=np.inf)


def replay_reference(ucrl2, spec, trace_i, optimization_horizon, kw, seed=SEED):
    """one loop's (s_t, a_t, s_tp1, reward bits) rows through the unmodified UCRL2Continuous, with MDPLoop.run's calls.
    Returns (agent, interaction times at which it ended its artificial episodes)."""
    ag = ucrl2.UCRL2Continuous(seed, spec, optimization_horizon, **kw)
    ag.before_start_interacting()
    ends = []
    for k in range(trace_i.shape[0]):
        s, a, sp, rbits = (int(x) for x in trace_i[k])
        r = float(np.int32(rbits).view(np.float32))
        ts = types.SimpleNamespace(observation=s)
        ts1 = types.SimpleNamespace(observation=sp, reward=r, last=lambda: False)
        ag.step_update(ts, a, ts1, k)
        if ag.is_episode_end(ts, a, ts1, k):
            ag.episode_end_update()
            ends.append(k + 2)  # our clock: counter 0 is the reset draw, row k is the step taken at time k + 1
    return ag, ends


def main():
    reference_models()
    ucrl2 = importlib.import_module("colosseum.agent.agents.infinite_horizon.ucrl2")
    out = {}
    for name, inst, kw in CASES:
        g = load_instance(inst)
        tb = MDPTables.from_golden(g)
        loops = orc.UCRL2Loops(host_tables(tb), N_LOOPS, N_STEPS, seed=SEED, **kw)
        trace = loops.steps(N_STEPS, trace=True)
        spec = mdp_spec(tb)
        ref = {k: [] for k in ("N", "P", "est_r", "var_r", "hold", "iteration", "episode", "delta", "Q", "ends")}
        for i in range(N_LOOPS):
            ag, ends = replay_reference(ucrl2, spec, trace[:, i], N_STEPS, kw)
            ref["ends"].append(np.asarray(ends, np.int64))
            ref["N"].append(ag.N.copy()); ref["P"].append(ag.P.copy())
            ref["est_r"].append(ag.estimated_rewards.copy()); ref["var_r"].append(ag.variance_proxy_reward.copy())
            ref["hold"].append(ag.estimated_holding_times.copy())
            ref["iteration"].append(ag.iteration); ref["episode"].append(ag.episode); ref["delta"].append(ag.delta)
            ref["Q"].append(np.asarray(ag.Q, np.float32).copy())
        out[f"{name}.trace"] = trace
        n_ends = max(len(e) for e in ref["ends"])
        ends = np.full((N_LOOPS, n_ends), -1, np.int64)
        for i, e in enumerate(ref["ends"]):
            ends[i, :len(e)] = e
        out[f"{name}.ref_ends"] = ends
        for k in ("N", "P", "est_r", "var_r", "hold", "Q"):
            out[f"{name}.ref_{k}"] = np.stack(ref[k])
        for k in ("iteration", "episode", "delta"):
            out[f"{name}.ref_{k}"] = np.asarray(ref[k])
        same_ends = all(list(ref["ends"][i]) == loops.episode_ends[i] for i in range(N_LOOPS))
        print(f"{name:22s} episodes/loop {[len(e) for e in ref['ends']]} episode ends identical: {same_ends}")
        for what, ours, r in (("N", loops.Nsas, out[f"{name}.ref_N"]), ("P", loops.P, out[f"{name}.ref_P"]),
                              ("est_r", loops.est_r, out[f"{name}.ref_est_r"]), ("var_r", loops.var_r, out[f"{name}.ref_var_r"]),
                              ("hold", loops.hold, out[f"{name}.ref_hold"]), ("iteration", loops.iteration, out[f"{name}.ref_iteration"]),
                              ("episode", loops.episode, out[f"{name}.ref_episode"]), ("delta", loops.delta, out[f"{name}.ref_delta"])):
            print(f"    {what:10s} identical: {np.array_equal(np.asarray(ours), np.asarray(r).astype(np.asarray(ours).dtype))}"
                  f"  (reference dtype {np.asarray(r).dtype})")
        print(f"    Q after the last episode end: max |oracle - reference| = {np.abs(loops.Q - out[f'{name}.ref_Q']).max():.2e}")
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ucrl2.npz"), **out)
    print("wrote tests/golden/ucrl2.npz")


if __name__ == "__main__":
    main()
