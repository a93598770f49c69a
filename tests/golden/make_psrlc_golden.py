"""Pins the oracle's PSRLContinuous loops (oracle.PSRLCLoops, orc_psrlc_steps) to the REFERENCE's agent class.

Run in the build container (needs /root/reference):  python tests/golden/make_psrlc_golden.py
The oracle runs N loops; at every artificial-episode end a deterministic planner (`plan_loop` below: the reference's
optimistic-sampling expressions with numpy draws seeded by (loop, episode), then the reference's in-place discounted VI
as restated by orc_discounted_gs_f32) produces the extended q-values.  Every loop's trajectory is then replayed through
the UNMODIFIED
    colosseum.agent.agents.infinite_horizon.posterior_sampling.PSRLContinuous
with MDPLoop.run's calls (before_start_interacting; step_update with the REAL action; is_episode_end;
episode_end_update -- the reference's own sampling and value iteration run, their output is not used), and
tests/golden/psrlc.npz stores the trace, the reference's episode ends, posteriors (N_NIG / N_N and Dirichlet
hyper-parameters), visit counts N, its (psi, omega, kappa, eta), and -- the deterministic half of optimistic_sampling --
the rows the reference itself produces for the under-visited (s, a) pairs when its randint draws are the oracle's z.
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
from make_ucrl2_golden import mdp_spec  # noqa: E402
from oracle import oracle as orc  # noqa: E402

CASES = [  # name, instance, agent kwargs (the reference's names)
    ("riverswim", "riverswimcontinuous_ergo0", dict(psi_weight=0.015, eta_weight=1e-9)),
    # (the reference's N_N.sample raises for every shape, conjugate_rewards.py:128-134: the sampling agents need N_NIG)
    ("frozenlake_eps", "frozenlakecontinuous_ergo0", dict(psi_weight=0.02, eta_weight=1e-8, epsilon_greedy=0.1,
                                                          rewards_prior_prms=[0.4, 2, 1.5, 3])),
    ("deepsea_plain", "deepsea10", dict(no_optimistic_sampling=True, transitions_prior_prms=[0.3])),
]
N_LOOPS, N_STEPS, SEED = 3, 3000, 31


def loop_kwargs(kw):
    return {k: v for k, v in kw.items() if k in ("epsilon_greedy", "rewards_prior_prms", "transitions_prior_prms",
                                                 "reward_prior_model")}


def parameters(tb, kw, T=N_STEPS):
    return orc.psrlc_parameters(tb.S, tb.A, T, **{k: v for k, v in kw.items() if k in
                                                  ("psi_weight", "omega_weight", "kappa_weight", "eta_weight",
                                                   "no_optimistic_sampling")})


def plan_loop(loops, i, prm):
    """extended q-values of loop i from its current posterior: optimistic_sampling + sample_R + discounted VI
    (posterior_sampling.py:347-376) with numpy draws seeded by (seed, loop, episode)"""
    S, A, psi, eta = loops.S, loops.A, loops.psi, prm["eta"]
    ep = int(loops.episode[i])
    rng = np.random.RandomState((loops.seed * 1000003 + (loops.env0 + int(i)) * 10007 + ep) % (2 ** 31))
    Nsum = loops.Nsa[i]
    cond = Nsum < eta
    Qt = np.zeros((psi, S, A, S), np.float32)
    for q in range(psi):
        if (~cond).any():
            hp = loops.dir_hyper[i][~cond]
            r = rng.standard_gamma(hp).astype(np.float32)
            Qt[q][~cond] = r / (1e-5 + r.sum(-1, keepdims=True))
        if cond.any():
            z = orc.psrlc_z(loops.seed, loops.env0 + int(i), ep, q, S)
            Qt[q][cond] = orc.psrlc_simple_rows(loops.Nsas[i], z)[cond]
    T = np.moveaxis(Qt, 0, 2).reshape((S, -1, S))
    h = loops.nig_hyper[i]
    if loops.reward_model == 1:
        R = rng.normal(h[..., 0], h[..., 1]).astype(np.float32)
    else:
        tau = rng.gamma(h[..., 2], 1 / h[..., 3]).astype(np.float32)
        R = rng.normal(h[..., 0], np.sqrt(1 / (h[..., 1] * tau))).astype(np.float32)
    R = np.tile(R, (1, psi))
    Q, _, _ = orc.discounted_gs_f32(np.ascontiguousarray(T), np.ascontiguousarray(R), gamma=0.99, eps=1e-3)
    return Q


def make_planner(prm):
    def planner(loops, idx):
        for i in idx:
            loops.Q[i] = plan_loop(loops, int(i), prm)
    return planner


class _ZRng:
    """stands in for the agent's RandomState while optimistic_sampling runs: randint returns the given states"""

    def __init__(self, zs):
        self.zs = list(zs)

    def randint(self, n):
        return self.zs.pop(0)


def replay_reference(psrl, spec, trace_i, psi, optimization_horizon, kw, seed=SEED):
    from colosseum.agent.mdp_models.bayesian_models import RewardsConjugateModel

    kw = dict(kw)
    if kw.get("reward_prior_model") == "N_N":
        kw["reward_prior_model"] = RewardsConjugateModel.N_N
    elif "rewards_prior_prms" in kw:
        kw["reward_prior_model"] = RewardsConjugateModel.N_NIG
    if "transitions_prior_prms" in kw:
        from colosseum.agent.mdp_models.bayesian_models import TransitionsConjugateModel

        kw["transitions_prior_model"] = TransitionsConjugateModel.M_DIR
    ag = psrl.PSRLContinuous(seed, spec, optimization_horizon, **kw)
    ag.before_start_interacting()
    ends = []
    for k in range(trace_i.shape[0]):
        s, a_ext, sp, rbits = (int(x) for x in trace_i[k])
        a = a_ext // psi  # extended_action_to_real (:449-452): MDPLoop hands the REAL action to step_update
        r = float(np.int32(rbits).view(np.float32))
        ts = types.SimpleNamespace(observation=s)
        ts1 = types.SimpleNamespace(observation=sp, reward=r, last=lambda: False)
        ag.step_update(ts, a, ts1, k)
        if ag.is_episode_end(ts, a, ts1, k):
            ag.episode_end_update()
            ends.append(k + 2)
    return ag, ends


def main():
    reference_models()
    psrl = importlib.import_module("colosseum.agent.agents.infinite_horizon.posterior_sampling")
    out = {}
    for name, inst, kw in CASES:
        g = load_instance(inst)
        tb = MDPTables.from_golden(g)
        prm = parameters(tb, kw)
        loops = orc.PSRLCLoops(host_tables(tb), N_LOOPS, prm["psi"], seed=SEED, planner=make_planner(prm), **loop_kwargs(kw))
        trace = loops.steps(N_STEPS, trace=True)
        ref = {k: [] for k in ("N", "nig", "dir", "ends", "prm", "simple", "cond")}
        for i in range(N_LOOPS):
            ag, ends = replay_reference(psrl, mdp_spec(tb), trace[:, i], prm["psi"], N_STEPS, kw)
            ref["ends"].append(np.asarray(ends, np.int64))
            ref["N"].append(ag.N.copy())
            ref["nig"].append(np.asarray(ag._mdp_model._rewards_model.hyper_params).copy())
            ref["dir"].append(np.asarray(ag._mdp_model._transitions_model.hyper_params).copy())
            ref["prm"].append([1 if ag.no_optimistic_sampling else ag.psi, ag.omega, ag.kappa,
                               0.0 if ag.no_optimistic_sampling else ag.eta])
            if not ag.no_optimistic_sampling:
                ep = int(loops.episode[i])
                zs = [orc.psrlc_z(SEED, i, ep, q, tb.S) for q in range(ag.psi)]
                ag._rng = _ZRng(zs)
                ag.optimistic_sampling()
                cond = ag.N.sum(-1) < ag.eta
                ref["cond"].append(cond)
                ref["simple"].append(np.where(cond[None, ..., None], ag.Q, 0).astype(np.float32))
                ours = np.stack([orc.psrlc_simple_rows(loops.Nsas[i], z) for z in zs])
                err = np.abs(np.where(cond[None, ..., None], ours - ag.Q, 0)).max()
                print(f"    loop {i}: {int(cond.sum())} under-visited pairs, simple-sampling rows max |oracle - reference| = {err:.1e}")
        out[f"{name}.trace"] = trace
        n_ends = max(len(e) for e in ref["ends"])
        ends = np.full((N_LOOPS, n_ends), -1, np.int64)
        for i, e in enumerate(ref["ends"]):
            ends[i, :len(e)] = e
        out[f"{name}.ref_ends"] = ends
        for k in ("N", "nig", "dir", "prm"):
            out[f"{name}.ref_{k}"] = np.stack([np.asarray(x) for x in ref[k]])
        if ref["simple"]:
            out[f"{name}.ref_simple"] = np.stack(ref["simple"])
            out[f"{name}.ref_cond"] = np.stack(ref["cond"])
        same_ends = all(list(ref["ends"][i]) == loops.episode_ends[i] for i in range(N_LOOPS))
        print(f"{name:20s} psi={prm['psi']} eta={prm['eta']:.3g} episodes/loop {[len(e) for e in ref['ends']]} "
              f"episode ends identical: {same_ends}; parameters identical: "
              f"{np.array_equal(out[f'{name}.ref_prm'][0], [prm['psi'], prm['omega'], prm['kappa'], prm['eta']])}")
        k = out[f"{name}.ref_nig"].shape[-1]
        for what, ours, r in (("N", loops.Nsas, out[f"{name}.ref_N"]), ("rew", loops.nig_hyper[..., :k], out[f"{name}.ref_nig"]),
                              ("dir", loops.dir_hyper, out[f"{name}.ref_dir"])):
            print(f"    {what:4s} identical: {np.array_equal(ours, np.asarray(r).astype(ours.dtype))} (reference dtype {np.asarray(r).dtype})")
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "psrlc.npz"), **out)
    print("wrote tests/golden/psrlc.npz")


if __name__ == "__main__":
    main()
