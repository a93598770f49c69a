"""GPU: the policy-induced Markov chain, its stationary distribution and the average-reward / regret indicators
(colosseum/mdp/utils/markov_chain.py:12-137, colosseum/experiment/indicators.py:9-45) against values recorded from
the unmodified reference (tests/golden/avg_reward.npz: the notebook's 0.99599 / 0.004008 / 0.5 among them)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_instance
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(GOLDEN, "avg_reward.npz"))


def test_average_reward_against_reference(g):
    import colosseum_b200.markov_chain as mc

    multi = 0
    for name in g["names"]:
        T, R = g[f"{name}_T"], g[f"{name}_R"]
        starts = list(zip(g[f"{name}_start_idx"].tolist(), g[f"{name}_start_prob"].tolist()))
        for k in ("opt", "worst", "rand"):
            pi = g[f"{name}_{k}_pi"]
            tps = mc.get_transition_probabilities(T, pi)
            assert tps.dtype == np.float32
            np.testing.assert_allclose(tps, g[f"{name}_{k}_tps"], rtol=2e-7, atol=1e-9)
            np.testing.assert_allclose(mc.get_average_rewards(R, pi), g[f"{name}_{k}_rs"], rtol=1e-6, atol=1e-9)
            Po, ro = orc.policy_chain(T, R, pi)
            x0 = np.zeros(len(T)); x0[g[f"{name}_start_idx"]] = g[f"{name}_start_prob"]
            sd = mc.get_stationary_distribution(tps, starts)
            n_cls = mc.get_stationary_distribution.last_classes
            multi += n_cls > 1
            # the start vector the reference's class rule yields (host logic), then the oracle's fp64 limit of it
            x0r, _ = mc.recurrent_class_weights(tps, starts)
            sd_o = orc.stationary_distribution_f64(Po, x0r)
            assert abs(sd.sum() - 1) < 1e-9 and sd.min() > -1e-15
            np.testing.assert_allclose(sd, sd_o, atol=2e-8)
            ar = mc.get_average_reward(T, R, pi, starts)
            assert abs(ar - float((sd_o * ro).sum())) < 1e-8
            # ... and the reference's own output (networkx recurrent classes + GTH, partly float32), including the
            # multichain policies where it assigns each start state's mass to the first class it can reach
            np.testing.assert_allclose(sd, g[f"{name}_{k}_sd"], atol=2e-5, err_msg=f"{name} {k} ({n_cls} classes)")
            assert abs(ar - float(g[f"{name}_{k}_ar"])) < 2e-6 * max(1.0, abs(ar)), (name, k)
    assert multi == 0  # every policy of this fixture is unichain; multichain ones: tests/golden/multichain.npz below


def test_notebook_average_rewards(g):
    """docs/_sources/mds/mdp-functionalities.ipynb:2369-2411"""
    import colosseum_b200.markov_chain as mc

    name = "doc_simplegrid4"
    T, R = g[f"{name}_T"], g[f"{name}_R"]
    starts = list(zip(g[f"{name}_start_idx"].tolist(), g[f"{name}_start_prob"].tolist()))
    for k, doc in (("opt", 0.9959919693746492), ("worst", 0.004008056365342812), ("rand", 0.5000000148429536)):
        assert abs(mc.get_average_reward(T, R, g[f"{name}_{k}_pi"], starts) - doc) < 2e-6


def test_episodic_regret_indicators():
    """experiment/indicators.py:9-45 on a golden episodic instance: GPU PE/VI vs the oracle"""
    import colosseum_b200.indicators as ind

    gi = load_instance("frozenlake4_epi")
    T, R, H = gi["T"], gi["R"], int(gi["H"])
    S, A = R.shape
    rs = np.random.RandomState(0)
    pol = rs.dirichlet(np.ones(A), size=(H, S)).astype(np.float32)
    start = np.zeros(S); start[gi["start_idx"]] = gi["start_prob"]
    reg, avg = ind.get_episodic_regrets_and_average_reward_at_time_zero(H, T, R, pol, start)
    _, Vp = orc.episodic_f64(H, T, R, pi=pol)
    _, Vs = orc.episodic_f64(H, T, R)
    np.testing.assert_allclose(reg, np.maximum(Vs[0] - Vp[0], 0), atol=2e-5)
    assert abs(avg - float((Vp[0] * start).sum())) < 2e-5
    np.testing.assert_allclose(ind.get_episodic_regret_at_time_zero(H, T, R, pol), Vs[0] - Vp[0], atol=2e-5)
    np.testing.assert_allclose(ind.get_episodic_regret_at_time_zero(H, T, R, pol, gi["vi_V"]), gi["vi_V"][0] - Vp[0], atol=2e-5)


def test_power_iteration_building_block(g):
    """colo_power_iteration_f64 (x <- M x / |M x| on the backup kernels, sparse on-chip and dense streaming paths) on
    well-mixing chains; and the documented failure mode on a nearly reducible one (why squaring is the default)"""
    import colosseum_b200.markov_chain as mc

    for name, k in (("deepsea10_prand", "rand"), ("frozenlake5", "rand"), ("doc_simplegrid4", "rand"), ("minigrid6", "opt")):
        tps = g[f"{name}_{k}_tps"]
        x0 = np.zeros(len(tps)); x0[g[f"{name}_start_idx"]] = g[f"{name}_start_prob"]
        x, it = mc.power_iteration(tps, x0, tol=1e-12)
        assert it > 10 and abs(x.sum() - 1) < 1e-12
        np.testing.assert_allclose(x, orc.stationary_distribution_f64(tps, x0), atol=1e-7)
    tps = g["doc_simplegrid4_opt_tps"]
    x0 = np.zeros(len(tps)); x0[g["doc_simplegrid4_start_idx"]] = g["doc_simplegrid4_start_prob"]
    x, it = mc.power_iteration(tps, x0, tol=1e-8)  # stops "converged" ...
    assert np.abs(x - orc.stationary_distribution_f64(tps, x0)).max() > 0.1  # ... far from the limit


def test_undiscounted_value_norm(g):
    """calculate_norm_average (hardness/measures/value_norm.py:64-93) vs a numpy fp64 restatement of the same two
    series (1e-9) and vs mdp.undiscounted_value_norm of the reference.  The reference takes the gain from a FLOAT32
    np.linalg.matrix_power(tps, 1000) (:64-66) and sums 1000 terms of (r - gain): a float32 error d in the gain shifts
    the bias by ~1000*d, which on the nearly reducible chain of SimpleGrid's optimal policy moves its own result by
    0.6 % (0.50131 vs 0.49848 in fp64).  Hence 2e-2 against the reference, and the fp64 restatement as the real bar."""
    import colosseum_b200.hardness as hd
    import colosseum_b200.markov_chain as mc

    for name in g["names"]:
        T, R, pi = g[f"{name}_T"], g[f"{name}_R"], g[f"{name}_opt_pi"]
        tps = mc.get_transition_probabilities(T, pi)
        ars = mc.get_average_rewards(R, pi)
        got = hd.calculate_norm_average(T, tps, ars)
        P = tps.astype(np.float64); r = ars.astype(np.float64)
        gain = np.linalg.matrix_power(P, 1000) @ r
        h = np.zeros(len(P)); v = r - gain
        for _ in range(1000):
            h += v; v = P @ v
        T64 = T.astype(np.float64)
        Eh = np.einsum("iaj,j->ia", T64, h)
        exp = np.sqrt(np.einsum("iaj,ja->ia", T64, (h.reshape(-1, 1) - Eh) ** 2)).max()
        assert abs(got - exp) < 1e-9 * max(1.0, exp), (name, got, exp)
        assert abs(got - float(g[f"{name}_undisc_norm"])) < 2e-2 * max(1.0, exp), (name, got, float(g[f"{name}_undisc_norm"]))


def test_multichain_policies_follow_the_reference_rule():
    """chains with several recurrent classes (tests/golden/multichain.npz, recorded from the reference's own
    get_stationary_distribution): each start state's whole mass goes to the first attracting component it can reach"""
    import colosseum_b200.markov_chain as mc

    g = np.load(os.path.join(GOLDEN, "multichain.npz"))
    for name in g["names"]:
        starts = list(zip(g[f"{name}_start_idx"].tolist(), g[f"{name}_start_prob"].tolist()))
        sd = mc.get_stationary_distribution(g[f"{name}_tps"], starts)
        assert mc.get_stationary_distribution.last_classes > 1
        np.testing.assert_allclose(sd, g[f"{name}_sd"], atol=2e-6, err_msg=str(name))
    with pytest.raises(TypeError):  # the reference needs the start distribution for a multichain policy
        mc.get_stationary_distribution(g["two_classes_tps"], None)


def test_batched_average_rewards(g):
    """colo_average_rewards_f64: the reference's optimal / worst / random policies of every fixture MDP plus 61 random
    deterministic policies, all in ONE batched solve, against the reference's recorded average rewards and the
    one-policy-at-a-time path; a multichain policy is flagged and resolved by the reference's class rule."""
    import colosseum_b200.markov_chain as mc

    rng = np.random.RandomState(0)
    for name in g["names"]:
        T, R = g[f"{name}_T"], g[f"{name}_R"]
        S, A = R.shape
        pis = [g[f"{name}_{k}_pi"] for k in ("opt", "worst", "rand")]
        pis += [np.eye(A, dtype=np.float32)[rng.randint(A, size=S)] for _ in range(61)]
        pis = np.stack(pis).astype(np.float32)
        start = int(g[f"{name}_start_idx"][0])
        ar = mc.get_average_reward_batched(T, R, pis, np.full(len(pis), start))
        for b, k in enumerate(("opt", "worst", "rand")):
            assert abs(ar[b] - float(g[f"{name}_{k}_ar"])) < 2e-6 * max(1.0, abs(ar[b])), (name, k)
        for b in range(3, len(pis), 7):
            one = mc.get_average_reward(T, R, pis[b], [(start, 1.0)])
            assert abs(ar[b] - one) < 1e-9, (name, b)
    # a policy with two closed classes: state 0 stays (reward 0), state 2 stays (reward 1), state 1 moves to 0
    T = np.zeros((3, 2, 3), np.float32)
    T[0, :, 0] = 1
    T[2, :, 2] = 1
    T[1, 0, 0] = 1
    T[1, 1, 2] = 1
    R = np.zeros((3, 2), np.float32)
    R[2] = 1
    pis = np.stack([np.eye(2, dtype=np.float32)[[0, 0, 0]], np.eye(2, dtype=np.float32)[[0, 1, 0]]])
    ar = mc.get_average_reward_batched(T, R, pis, np.array([1, 1]))
    assert mc.get_average_reward_batched.last_multichain == 2
    assert abs(ar[0] - 0.0) < 1e-12 and abs(ar[1] - 1.0) < 1e-12  # from state 1: into class {0} / into class {2}
    assert abs(mc.get_average_reward_batched(T, R, pis, np.array([2, 0]))[0] - 1.0) < 1e-12


def test_continuous_regret_of_every_loop():
    """BatchedMDPLoop.run(regret_for="all") on a continuous MDP: every loop's greedy policy through the batched
    stationary-distribution solve == the per-loop indicator, for the three continuous agents."""
    import colosseum_b200.agent_loop as al
    from colosseum_b200.tables import MDPTables

    gi = load_instance("riverswimcontinuous_ergo0")
    tb = MDPTables.from_golden(gi)
    T, R = np.asarray(gi["T"], np.float32), np.asarray(gi["R"], np.float32)
    for make in (lambda: al.QLearningContinuous(3, tb, 4000, n_loops=40),
                 lambda: al.UCRL2Continuous(3, tb, 4001, alpha_r=0.1, alpha_p=0.05, n_loops=40),
                 lambda: al.PSRLContinuous(3, tb, 4001, psi_weight=0.015, eta_weight=1e-9, n_loops=40)):
        ag = make()
        loop = al.BatchedMDPLoop(ag, T, R)
        logs = loop.run(4000, log_every=2000, regret_for="all")
        assert len(logs) == 2 and logs[-1]["regret"].shape == (40,)
        import colosseum_b200.dynamic_programming as dp
        import colosseum_b200.markov_chain as mc

        Qo, _ = dp.discounted_value_iteration(T, R)
        opt = mc.get_average_reward(T, R, dp.get_policy_from_q_values(Qo, True))
        acts, states = loop._greedy_actions.cpu().numpy(), ag.state.cpu().numpy()
        for i in (0, 7, 39):  # the same policy through the one-policy-at-a-time indicator
            ar = mc.get_average_reward(T, R, np.eye(tb.A, dtype=np.float32)[acts[i]], [(int(states[i]), 1.0)])
            r = opt - ar
            want = 0.0 if (np.isclose(r, 0.0, atol=1e-3) or r < 0) else r
            assert abs(logs[-1]["regret"][i] - want) < 2e-6, i
        assert (logs[-1]["regret"] >= 0).all() and (logs[-1]["cumulative_regret"] >= logs[0]["cumulative_regret"]).all()
        assert np.isfinite(logs[-1]["cumulative_expected_reward"]).all()
