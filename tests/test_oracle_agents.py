"""CPU: the oracle's Q-learning loops (oracle.QLearningLoops) are pinned to the reference's model classes through
tests/golden/qlearning.npz (made by tests/golden/make_qlearning_golden.py, which replays the oracle's transitions
through the unmodified `QValuesModel` / `_QValuesModel`)."""
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN, load_instance
from colosseum_b200.tables import MDPTables
from oracle import oracle as orc

sys.path.insert(0, GOLDEN)
from make_qlearning_golden import CASES, N_LOOPS, N_STEPS, SEED, host_tables  # noqa: E402


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, "qlearning.npz"))


@pytest.mark.parametrize("name,inst,kw", CASES, ids=[c[0] for c in CASES])
def test_oracle_loops_match_reference_models(gold, name, inst, kw):
    tb = MDPTables.from_golden(load_instance(inst))
    loops = orc.QLearningLoops(host_tables(tb), N_LOOPS, seed=SEED, **kw)
    assert np.array_equal(loops.state, gold[f"{name}.start"])
    trace = loops.steps(N_STEPS, trace=True)
    assert np.array_equal(trace, gold[f"{name}.trace"])
    ours = {"N": loops.cnt, "Q": loops.Q, "V": loops.V}
    for f in ("Q_main", "mu", "sigma", "beta"):
        if hasattr(loops, f):
            ours[f] = getattr(loops, f)
    for f, v in ours.items():
        ref = gold[f"{name}.ref_{f}"]
        if tb.H > 0:
            # episodic model: every operation restated with numpy's own promotion -> bit-identical tables
            assert np.array_equal(v, ref.astype(v.dtype)), f
        else:
            # continuous model: under numpy >= 2 the reference's `np.zeros(float32) + np.float64` tables silently
            # become float64; the restatement keeps the declared float32 -> agreement to float32 rounding
            np.testing.assert_allclose(v, ref, rtol=1e-6, atol=0, err_msg=f)
    # the trajectory is a real interaction: actions cover the action set, episodes end where they should
    assert set(np.unique(trace[..., 1])) <= set(range(tb.A))
    if tb.H > 0:
        assert int(loops.n_episodes.min()) == N_STEPS // tb.H
        assert ((trace[..., 2] == -1).sum(0) == N_STEPS // tb.H).all()


def test_chunked_steps_equal_one_call():
    tb = MDPTables.from_golden(load_instance("frozenlake4_epi"))
    kw = dict(optimization_horizon=1000, p=0.05, c_1=0.5, c_2=0.5, UCB_type="bernstein")
    a = orc.QLearningLoops(host_tables(tb), 5, seed=3, **kw)
    b = orc.QLearningLoops(host_tables(tb), 5, seed=3, **kw)
    ta = a.steps(200, trace=True)
    tb_ = np.concatenate([b.steps(70, trace=True), b.steps(130, trace=True)])
    assert np.array_equal(ta, tb_) and np.array_equal(a.Q, b.Q) and np.array_equal(a.cum_reward, b.cum_reward)


def test_oracle_psrl_updates_match_reference_conjugate_models():
    """oracle.PSRLLoops' posterior updates == the reference's BayesianMDPModel.step_update (N_NIG + M_DIR), bit for
    bit, on the committed trajectories (tests/golden/make_psrl_golden.py)."""
    from make_psrl_golden import CASES as PCASES, N_EPISODES, N_LOOPS as PN, SEED as PSEED, optimal_q

    gold = np.load(os.path.join(GOLDEN, "psrl.npz"))
    for name, inst, kw in PCASES:
        g = load_instance(inst)
        tb = MDPTables.from_golden(g)
        loops = orc.PSRLLoops(host_tables(tb), PN, seed=PSEED, **kw)
        loops.set_q(optimal_q(g, tb))
        trace = loops.steps(N_EPISODES * tb.H, trace=True)
        assert np.array_equal(trace, gold[f"{name}.trace"]), name
        k = gold[f"{name}.ref_nig"].shape[-1]  # 4 parameters for N_NIG, 2 for N_N
        assert np.array_equal(loops.nig_hyper[..., :k], gold[f"{name}.ref_nig"]), name
        assert np.array_equal(loops.dir_hyper, gold[f"{name}.ref_dir"]), name
        assert (loops.n_episodes == N_EPISODES).all()
        # transitions into the terminal observation are not counted (bayesian_model.py:89-92)
        assert np.isclose((loops.dir_hyper - loops.dir_hyper.min()).sum(), PN * N_EPISODES * (tb.H - 1), rtol=1e-3)


def test_oracle_ucrl2_loops_match_reference_agent():
    """oracle.UCRL2Loops == the reference's UCRL2Continuous class replayed on the committed trajectories
    (tests/golden/make_ucrl2_golden.py): same artificial-episode ends, bit-identical model tables (N, P, estimated
    rewards, variance proxies, holding times, iteration / episode counters, delta), and the extended-VI Q within its
    stopping tolerance."""
    from make_ucrl2_golden import CASES as UCASES, N_LOOPS as UN, N_STEPS as UT, SEED as USEED

    gold = np.load(os.path.join(GOLDEN, "ucrl2.npz"))
    for name, inst, kw in UCASES:
        tb = MDPTables.from_golden(load_instance(inst))
        loops = orc.UCRL2Loops(host_tables(tb), UN, UT, seed=USEED, **kw)
        trace = loops.steps(UT, trace=True)
        assert np.array_equal(trace, gold[f"{name}.trace"]), name
        ends = gold[f"{name}.ref_ends"]
        for i in range(UN):
            assert loops.episode_ends[i] == [int(x) for x in ends[i] if x >= 0], (name, i)
        for f, v in (("N", loops.Nsas), ("P", loops.P), ("est_r", loops.est_r), ("var_r", loops.var_r),
                     ("hold", loops.hold), ("iteration", loops.iteration), ("episode", loops.episode),
                     ("delta", loops.delta)):
            ref = gold[f"{name}.ref_{f}"]
            assert np.array_equal(v, ref.astype(v.dtype)), (name, f)
        assert np.abs(loops.Q - gold[f"{name}.ref_Q"]).max() < 2e-3, name  # extended VI stops at eps = 1e-3
        assert len(ends[0]) > 100


def test_oracle_psrl_continuous_loops_match_reference_agent():
    """oracle.PSRLCLoops == the reference's PSRLContinuous class replayed on the committed trajectories
    (tests/golden/make_psrlc_golden.py): same artificial-episode ends, bit-identical visit counts and posteriors, the
    reference's own (psi, omega, kappa, eta), and the reference's own "simple sampling" rows of optimistic_sampling."""
    import make_psrlc_golden as mk

    gold = np.load(os.path.join(GOLDEN, "psrlc.npz"))
    for name, inst, kw in mk.CASES:
        tb = MDPTables.from_golden(load_instance(inst))
        prm = mk.parameters(tb, kw)
        assert np.array_equal(gold[f"{name}.ref_prm"][0], [prm["psi"], prm["omega"], prm["kappa"], prm["eta"]]), name
        loops = orc.PSRLCLoops(host_tables(tb), mk.N_LOOPS, prm["psi"], seed=mk.SEED, planner=mk.make_planner(prm),
                               **mk.loop_kwargs(kw))
        trace = loops.steps(mk.N_STEPS, trace=True)
        assert np.array_equal(trace, gold[f"{name}.trace"]), name
        ends = gold[f"{name}.ref_ends"]
        for i in range(mk.N_LOOPS):
            assert loops.episode_ends[i] == [int(x) for x in ends[i] if x >= 0], (name, i)
        k = gold[f"{name}.ref_nig"].shape[-1]
        for f, v, ref in (("N", loops.Nsas, gold[f"{name}.ref_N"]), ("rew", loops.nig_hyper[..., :k], gold[f"{name}.ref_nig"]),
                          ("dir", loops.dir_hyper, gold[f"{name}.ref_dir"])):
            assert np.array_equal(v, ref.astype(v.dtype)), (name, f)
        if f"{name}.ref_simple" in gold.files:
            for i in range(mk.N_LOOPS):
                cond = gold[f"{name}.ref_cond"][i]
                assert cond.any()
                ep = int(loops.episode[i])
                for q in range(prm["psi"]):
                    z = orc.psrlc_z(mk.SEED, i, ep, q, tb.S)
                    ours = orc.psrlc_simple_rows(loops.Nsas[i], z)
                    np.testing.assert_allclose(ours[cond], gold[f"{name}.ref_simple"][i, q][cond], atol=1e-7, rtol=0)


def test_oracle_boltzmann_matches_reference_actor():
    """orc_boltzmann_action == the reference's QValuesActor.select_action with a temperature schedule, on the committed
    (q row, temperature, uniform) triples recorded from the unmodified class (tests/golden/make_actor_golden.py)."""
    gold = np.load(os.path.join(GOLDEN, "actor.npz"))
    names = sorted({k.split(".")[0] for k in gold.files})
    assert len(names) == 3
    for nm in names:
        Q, st, u, temp, act = (gold[f"{nm}.{k}"] for k in ("Q", "states", "u", "temp", "action"))
        ours = np.array([orc.boltzmann_action(Q[s], t, x) for s, t, x in zip(st, temp, u)])
        assert np.array_equal(ours, act), nm


def test_oracle_exploration_schedules():
    """epsilon / temperature as functions of the interaction counter: a schedule that is constant equals the constant;
    a step schedule equals its two constants applied in turn"""
    tb = MDPTables.from_golden(load_instance("frozenlakecontinuous_ergo0"))
    kw = dict(optimization_horizon=5000, min_at=0.02)
    a = orc.QLearningLoops(host_tables(tb), 8, seed=4, epsilon_greedy=0.3, **kw)
    b = orc.QLearningLoops(host_tables(tb), 8, seed=4, epsilon_greedy=lambda t: 0.3, **kw)
    assert np.array_equal(a.steps(300, trace=True), np.concatenate([b.steps(100, trace=True), b.steps(200, trace=True)]))
    c = orc.QLearningLoops(host_tables(tb), 8, seed=4, boltzmann_temperature=0.7, **kw)
    d = orc.QLearningLoops(host_tables(tb), 8, seed=4, boltzmann_temperature=lambda t: 0.7, **kw)
    tc = c.steps(300, trace=True)
    assert np.array_equal(tc, d.steps(300, trace=True)) and not np.array_equal(tc, a.steps(300, trace=True)[:300])
    assert len(np.unique(tc[..., 1])) == tb.A
    # a step schedule == the two constants applied in turn (interaction counts 1..50 / 51..)
    e = orc.QLearningLoops(host_tables(tb), 64, seed=4, epsilon_greedy=lambda t: 0.2 if t <= 50 else 0.9, **kw)
    g = orc.QLearningLoops(host_tables(tb), 64, seed=4, epsilon_greedy=0.2, **kw)
    te = e.steps(120, trace=True)
    tg1 = g.steps(50, trace=True)
    g._explore = (0.9, None)
    assert np.array_equal(te, np.concatenate([tg1, g.steps(70, trace=True)]))
