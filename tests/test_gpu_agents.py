"""GPU: the batched agent/MDP loops (csrc/agents.cu through colosseum_b200.agent_loop) against the oracle, bit for bit
(trajectories and every agent table), and end-to-end learning sanity through BatchedMDPLoop."""
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN, load_instance
from colosseum_b200.tables import MDPTables
from oracle import oracle as orc

sys.path.insert(0, GOLDEN)
from make_qlearning_golden import CASES, host_tables  # noqa: E402

pytestmark = pytest.mark.gpu


def make_agents(tb, kw, n_loops, seed):
    import colosseum_b200.agent_loop as al

    kw = dict(kw)
    T = kw.pop("optimization_horizon")
    if tb.H > 0:
        return al.QLearningEpisodic(seed, tb, T, n_loops=n_loops, **kw)
    return al.QLearningContinuous(seed, tb, T, n_loops=n_loops, **kw)


@pytest.mark.parametrize("name,inst,kw", CASES, ids=[c[0] for c in CASES])
def test_kernel_equals_oracle_bit_for_bit(name, inst, kw):
    tb = MDPTables.from_golden(load_instance(inst))
    N, n_steps, seed = 333, 250, 21  # ragged last CTA
    dev = make_agents(tb, kw, N, seed)
    cpu = orc.QLearningLoops(host_tables(tb), N, seed=seed, **kw)
    assert np.array_equal(dev.state.cpu().numpy(), cpu.state)
    tr_d = np.concatenate([dev.steps(100, trace=True).cpu().numpy(), dev.steps(n_steps - 100, trace=True).cpu().numpy()])
    tr_c = cpu.steps(n_steps, trace=True)
    assert np.array_equal(tr_d, tr_c)
    pairs = [(dev.N, cpu.cnt), (dev.Q, cpu.Q), (dev.V, cpu.V), (dev.state, cpu.state), (dev.h, cpu.h),
             (dev.cumulative_reward, cpu.cum_reward), (dev.n_episodes, cpu.n_episodes)]
    for f in ("Q_main", "mu", "sigma", "beta"):
        if hasattr(cpu, f):
            pairs.append((getattr(dev, f), getattr(cpu, f)))
    for d, c in pairs:
        assert np.array_equal(d.cpu().numpy(), c)


def test_golden_trace_is_reproduced_on_the_gpu():
    """the committed trace (whose replay through the REFERENCE model classes gave the golden tables) comes out of the
    kernel unchanged, and so do the reference's tables"""
    from make_qlearning_golden import N_LOOPS, N_STEPS, SEED

    gold = np.load(os.path.join(GOLDEN, "qlearning.npz"))
    for name, inst, kw in CASES:
        tb = MDPTables.from_golden(load_instance(inst))
        dev = make_agents(tb, kw, N_LOOPS, SEED)
        tr = dev.steps(N_STEPS, trace=True).cpu().numpy()
        assert np.array_equal(tr, gold[f"{name}.trace"]), name
        if tb.H > 0:
            assert np.array_equal(dev.Q.cpu().numpy(), gold[f"{name}.ref_Q"]), name
            assert np.array_equal(dev.N.cpu().numpy(), gold[f"{name}.ref_N"]), name
        else:
            np.testing.assert_allclose(dev.Q.cpu().numpy(), gold[f"{name}.ref_Q"], rtol=1e-6)


def test_batched_loop_learns_and_logs_regret():
    """BatchedMDPLoop on C1 (RiverSwimEpisodic size 5): 64 seeds, 40,000 steps.  The log records have the MDPLoop
    shape, the expected regret of the greedy policies is a valid regret (0 <= regret <= V*[0, s0] / H), and the
    cumulative reward beats a uniformly random agent's (UCB Q-learning with the paper's constants is still exploring
    after 40,000 steps -- the CPU restatement shows the same -- so the regret itself is not required to fall)."""
    import colosseum_b200.agent_loop as al

    g = load_instance("c1_riverswim_epi")
    tb = MDPTables.from_golden(g)
    T_steps, N = 40000, 64
    agents = al.QLearningEpisodic(0, tb, T_steps, p=0.05, c_1=0.05, c_2=0.05, min_at=0.0, UCB_type="bernstein",
                                  n_loops=N)
    loop = al.BatchedMDPLoop(agents, T=np.asarray(g["T"], np.float32), R=np.asarray(g["R"], np.float32))
    logs = loop.run(T_steps, log_every=10000, regret_for=range(8))
    assert [r["steps"] for r in logs] == [10000, 20000, 30000, 40000]
    assert (logs[-1]["n_episodes"] == T_steps // tb.H).all()
    assert (np.diff([r["cumulative_regret"] for r in logs], axis=0) >= 0).all()
    from colosseum_b200.dynamic_programming import episodic_value_iteration

    v_star = episodic_value_iteration(tb.H, np.asarray(g["T"], np.float32), np.asarray(g["R"], np.float32))[1]
    bound = float(v_star[0][int(tb.start_idx[0])]) / tb.H
    for r in logs:
        assert (r["regret"] >= 0).all() and (r["regret"] <= bound + 1e-6).all()
    rnd = al.QLearningEpisodic(0, tb, T_steps, p=0.05, c_1=0.05, epsilon_greedy=1.0, n_loops=N)
    rnd.steps(T_steps)
    assert logs[-1]["cumulative_reward"].mean() > rnd.cumulative_reward.mean().item()


def test_psrl_steps_equal_oracle_and_reference_posteriors():
    """colo_psrl_episodic_steps against the oracle with the SAME Q per episode (bit for bit: trace, NIG and Dirichlet
    parameters), and the committed reference posteriors reproduced on the GPU."""
    import torch

    import colosseum_b200.agent_loop as al
    from make_psrl_golden import CASES as PCASES, N_EPISODES, N_LOOPS as PN, SEED as PSEED, optimal_q

    gold = np.load(os.path.join(GOLDEN, "psrl.npz"))
    for name, inst, kw in PCASES:
        g = load_instance(inst)
        tb = MDPTables.from_golden(g)
        # (1) golden: act on the true optimal Q, no resampling -> the reference's posteriors
        dev = al.PSRLEpisodic(PSEED, tb, 10 ** 5, n_loops=PN, **kw)
        dev.episode_end_update = lambda: None
        dev.Q.copy_(torch.from_numpy(optimal_q(g, tb)).cuda()[None].expand_as(dev.Q))
        tr = dev.steps(N_EPISODES * tb.H, trace=True).cpu().numpy()
        assert np.array_equal(tr, gold[f"{name}.trace"]), name
        k = gold[f"{name}.ref_nig"].shape[-1]  # 4 parameters for N_NIG, 2 for N_N
        assert np.array_equal(dev.nig_hyper.cpu().numpy()[..., :k], gold[f"{name}.ref_nig"]), name
        assert np.array_equal(dev.dir_hyper.cpu().numpy(), gold[f"{name}.ref_dir"]), name
        # (2) the real agent (resampling every episode): the oracle follows with the GPU's sampled-model Q
        N = 200
        dev = al.PSRLEpisodic(3, tb, 10 ** 5, n_loops=N, **kw)
        cpu = orc.PSRLLoops(host_tables(tb), N, seed=3, **kw)
        assert np.array_equal(dev.state.cpu().numpy(), cpu.state)
        for ep in range(6):
            cpu.set_q(dev.Q.cpu().numpy())
            td = dev.steps(tb.H, trace=True).cpu().numpy()
            tc = cpu.steps(tb.H, trace=True)
            assert np.array_equal(td, tc), (name, ep)
        assert np.array_equal(dev.nig_hyper.cpu().numpy(), cpu.nig_hyper)
        assert np.array_equal(dev.dir_hyper.cpu().numpy(), cpu.dir_hyper)
        assert np.array_equal(dev.cumulative_reward.cpu().numpy(), cpu.cum_reward)
        # sampled models are sub-stochastic by the reference's 1e-5 in the denominator (conjugate_transitions.py:53)
        rs = dev.T_sample.sum(-1)
        assert float(rs.max()) <= 1 + 1e-5 and float(rs.min()) > 0.5 and bool((dev.T_sample >= 0).all())


def test_nig_sampler_distribution():
    """colo_sample_nig_rewards: the sampled mean of a Normal-Inverse-Gamma(mu, lambda, alpha, beta) posterior is
    Student-t with 2*alpha degrees of freedom, location mu, scale sqrt(beta / (alpha*lambda)) -- KS test vs scipy."""
    import scipy.stats
    import torch

    from colosseum_b200 import _cabi

    n = 200000
    for mu, lam, alpha, beta in ((0.5, 3.0, 2.5, 1.2), (-1.0, 1.0, 0.7, 0.3), (2.0, 50.0, 30.0, 4.0)):
        hyper = torch.tensor([mu, lam, alpha, beta], dtype=torch.float32, device="cuda").repeat(n, 1).contiguous()
        out = torch.empty(n, dtype=torch.float32, device="cuda")
        rc = _cabi.lib().colo_sample_nig_rewards(hyper.data_ptr(), n, 0, 11, 0, out.data_ptr(), _cabi.current_stream())
        assert rc == 0
        x = out.cpu().numpy().astype(np.float64)
        ks = scipy.stats.kstest(x, scipy.stats.t(df=2 * alpha, loc=mu, scale=np.sqrt(beta / (alpha * lam))).cdf)
        assert ks.pvalue > 1e-3, (mu, lam, alpha, beta, ks)
        # another draw counter gives another sample
        out2 = torch.empty_like(out)
        _cabi.lib().colo_sample_nig_rewards(hyper.data_ptr(), n, 0, 11, 1, out2.data_ptr(), _cabi.current_stream())
        assert not torch.equal(out, out2)


def test_psrl_learns_river_swim():
    """PSRLEpisodic on C1 through BatchedMDPLoop: after 5,000 steps the MAP-policy regret of every logged loop is far
    below the first tick's, and the cumulative reward beats the random agent's by a wide margin."""
    import colosseum_b200.agent_loop as al

    g = load_instance("c1_riverswim_epi")
    tb = MDPTables.from_golden(g)
    agents = al.PSRLEpisodic(0, tb, 5000, n_loops=64)
    loop = al.BatchedMDPLoop(agents, T=np.asarray(g["T"], np.float32), R=np.asarray(g["R"], np.float32))
    logs = loop.run(5000, log_every=500, regret_for=range(6))
    assert logs[-1]["regret"].mean() < 0.25 * max(logs[0]["regret"].mean(), 1e-3) + 2e-3
    rnd = al.QLearningEpisodic(0, tb, 5000, p=0.05, c_1=0.05, epsilon_greedy=1.0, n_loops=64)
    rnd.steps(5000)
    assert logs[-1]["cumulative_reward"].mean() > 1.15 * rnd.cumulative_reward.mean().item()


def test_agent_checkpoint_resume_is_bit_exact():
    import colosseum_b200.agent_loop as al

    tb = MDPTables.from_golden(load_instance("frozenlake4_epi"))
    for make in (lambda s: al.QLearningEpisodic(s, tb, 10 ** 4, p=0.05, c_1=0.4, c_2=0.4, UCB_type="bernstein", n_loops=100),
                 lambda s: al.PSRLEpisodic(s, tb, 10 ** 4, n_loops=100)):
        a = make(7)
        a.steps(5 * tb.H + 3)
        ck = al.agents_state_dict(a)
        tr_a = a.steps(4 * tb.H, trace=True).cpu().numpy()
        b = make(7)
        al.agents_load_state_dict(b, ck)
        tr_b = b.steps(4 * tb.H, trace=True).cpu().numpy()
        assert np.array_equal(tr_a, tr_b)
        assert np.array_equal(a.cumulative_reward.cpu().numpy(), b.cumulative_reward.cpu().numpy())


def test_batched_regret_of_all_loops_matches_per_loop_indicator():
    """regret_for='all' (one colo_episodic_policies_f32 launch per tick) against the per-loop indicator path, on
    policies without argmax ties (ties are broken by different random streams), and against the oracle's PE"""
    import torch

    import colosseum_b200.agent_loop as al
    from colosseum_b200 import _cabi

    g = load_instance("taxi_epi")
    tb = MDPTables.from_golden(g)
    T, R = np.asarray(g["T"], np.float32), np.asarray(g["R"], np.float32)
    N = 40
    ag = al.QLearningEpisodic(1, tb, 10 ** 5, p=0.05, c_1=0.3, n_loops=N)
    ag.steps(500)
    torch.manual_seed(0)
    ag.Q.add_(torch.rand_like(ag.Q))  # no exact ties (float32 spacing at Q = H is 2e-6): same greedy actions on both paths
    loop = al.BatchedMDPLoop(ag, T=T, R=R)
    all_reg = loop._expected_regret_all(chunk=16)
    some = loop._expected_regret([0, 5, 17, 39])
    # both paths evaluate in float32 with different summation orders; the regret is a difference of values ~20x larger
    np.testing.assert_allclose(all_reg[[0, 5, 17, 39]], some, rtol=5e-4, atol=1e-5)
    # the shared-MDP policy evaluation itself against the oracle
    rs = np.random.RandomState(0)
    pol = rs.dirichlet(np.ones(tb.A), size=(3, tb.H, tb.S)).astype(np.float32)
    pd = torch.from_numpy(pol).cuda()
    Q = torch.empty((3, tb.H + 1, tb.S, tb.A), dtype=torch.float64, device="cuda")
    V = torch.empty((3, tb.H + 1, tb.S), dtype=torch.float64, device="cuda")
    Td, Rd = torch.from_numpy(T).cuda(), torch.from_numpy(R).cuda()
    rc = _cabi.lib().colo_episodic_policies_f64acc(Td.data_ptr(), Rd.data_ptr(), pd.data_ptr(), 3, tb.S, tb.A, tb.H,
                                                   Q.data_ptr(), V.data_ptr(), _cabi.current_stream())
    assert rc == 0
    for b in range(3):
        Qo, Vo = orc.episodic_f64(tb.H, T, R, pi=pol[b])
        np.testing.assert_allclose(V[b].cpu().numpy(), Vo, rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(Q[b].cpu().numpy(), Qo, rtol=1e-9, atol=1e-12)
    # and through run(): every loop logged
    logs = loop.run(200, log_every=100, regret_for="all")
    assert logs[-1]["regret"].shape == (N,) and (logs[-1]["cumulative_regret"] >= logs[0]["cumulative_regret"]).all()
    # MDPLoop's normalised indicators: baselines from the oracle, identities between the log fields
    Vo = orc.episodic_f64(tb.H, T, R)[1][0]
    Vw = -orc.episodic_f64(tb.H, T, -R)[1][0]
    Vr = orc.episodic_f64(tb.H, T, R, pi=np.full((tb.H, tb.S, tb.A), 1.0 / tb.A))[1][0]
    p0 = np.zeros(tb.S)
    pr = np.diff(np.concatenate([[0.0], tb.start_cum]))
    np.add.at(p0, tb.start_idx, pr / pr.sum())
    base = loop.episodic_baselines()
    for k, v in (("optimal", Vo), ("worst", Vw), ("random", Vr)):
        assert abs(base[k] - (v * p0).sum() / tb.H) < 1e-9, k
    last = logs[-1]
    span = base["optimal"] - base["worst"]
    np.testing.assert_allclose(last["normalized_cumulative_regret"], last["cumulative_regret"] / span)
    assert abs(last["optimal_cumulative_expected_reward"] - base["optimal"] * last["steps"]) < 1e-9
    # expected reward of the agent + its regret = the optimal expected reward, tick by tick
    np.testing.assert_allclose(last["cumulative_expected_reward"] + last["cumulative_regret"],
                               last["optimal_cumulative_expected_reward"], rtol=1e-4)
    assert last["steps_per_second"] > 0


def test_loops_sharded_over_ranks_equal_the_unsharded_batch():
    """env_offset: loops [32, 64) run as their own batch (another GPU's shard) walk exactly the trajectories they walk
    inside the batch of 64 -- Q-learning and PSRL (whose posterior samples are keyed by the global row index)."""
    import colosseum_b200.agent_loop as al

    tb = MDPTables.from_golden(load_instance("frozenlake4_epi"))
    for make in (lambda n, off: al.QLearningEpisodic(9, tb, 10 ** 4, p=0.05, c_1=0.4, n_loops=n, env_offset=off),
                 lambda n, off: al.PSRLEpisodic(9, tb, 10 ** 4, n_loops=n, env_offset=off)):
        whole, lo, hi = make(64, 0), make(32, 0), make(32, 32)
        n = 5 * tb.H
        tw = whole.steps(n, trace=True).cpu().numpy()
        tl, th = lo.steps(n, trace=True).cpu().numpy(), hi.steps(n, trace=True).cpu().numpy()
        assert np.array_equal(tw[:, :32], tl) and np.array_equal(tw[:, 32:], th)
        assert np.array_equal(whole.cumulative_reward.cpu().numpy()[32:], hi.cumulative_reward.cpu().numpy())


EXPLORE = [("eps_schedule", dict(epsilon_greedy=lambda t: 1.0 / (1.0 + 0.01 * t))),
           ("boltzmann_const", dict(boltzmann_temperature=0.8)),
           ("boltzmann_schedule_eps", dict(boltzmann_temperature=lambda t: 0.2 + 0.001 * t, epsilon_greedy=0.05))]


@pytest.mark.parametrize("tag,ex", EXPLORE, ids=[e[0] for e in EXPLORE])
def test_exploration_kernel_equals_oracle_bit_for_bit(tag, ex):
    """QValuesActor's full exploration set on the device -- epsilon / temperature as constants or as functions of the
    interaction counter, Boltzmann action draws (pinned to the reference actor by tests/golden/actor.npz) -- for the
    Q-learning agents (episodic + continuous), PSRLEpisodic, UCRL2Continuous and PSRLContinuous: trajectories and tables
    equal the oracle's bit for bit, across launches that split the schedule."""
    import torch

    import colosseum_b200.agent_loop as al

    N, seed = 77, 13
    # Q-learning, episodic and continuous
    for inst, kw in (("frozenlake4_epi", dict(optimization_horizon=3000, p=0.05, c_1=0.4, c_2=0.9, min_at=0.05, UCB_type="bernstein")),
                     ("frozenlakecontinuous_ergo0", dict(optimization_horizon=5000, min_at=0.02))):
        tb = MDPTables.from_golden(load_instance(inst))
        dev = make_agents(tb, {**kw, **ex}, N, seed)
        cpu = orc.QLearningLoops(host_tables(tb), N, seed=seed, **kw, **ex)
        tr_d = torch.cat([dev.steps(90, trace=True), dev.steps(160, trace=True)]).cpu().numpy()
        assert np.array_equal(tr_d, cpu.steps(250, trace=True)), inst
        assert np.array_equal(dev.Q.cpu().numpy(), cpu.Q) and np.array_equal(dev.N.cpu().numpy(), cpu.cnt)
        assert len(np.unique(tr_d[..., 1])) == tb.A
    # PSRLEpisodic between posterior samples (the oracle has no sampler: both act on the device's sampled Q)
    tb = MDPTables.from_golden(load_instance("frozenlake4_epi"))
    dev = al.PSRLEpisodic(seed, tb, 2000, n_loops=N, **ex)
    cpu = orc.PSRLLoops(host_tables(tb), N, seed=seed, **ex)
    cpu.set_q(dev.Q.cpu().numpy())
    assert np.array_equal(dev.steps(tb.H - 1, trace=True).cpu().numpy(), cpu.steps(tb.H - 1, trace=True))
    assert np.array_equal(dev.nig_hyper.cpu().numpy(), cpu.nig_hyper)
    # the continuous model-based agents, with the planners replaced by one fixed random q-table on both sides
    tb = MDPTables.from_golden(load_instance("riverswimcontinuous_ergo0"))
    rng = np.random.RandomState(1)
    Qfix = rng.rand(tb.S, tb.A).astype(np.float32)

    def dev_planner_u(ag, idx, br, bp):
        ag.Q[idx.long()] = torch.from_numpy(Qfix).cuda()

    du = al.UCRL2Continuous(seed, tb, 801, n_loops=N, planner=dev_planner_u, **ex)
    cu = orc.UCRL2Loops(host_tables(tb), N, 801, seed=seed, planner=lambda i, e, *a: (0.0, Qfix, np.zeros(tb.S, np.float32)), **ex)
    assert np.array_equal(torch.cat([du.steps(300, trace=True), du.steps(500, trace=True)]).cpu().numpy(), cu.steps(800, trace=True))
    assert np.array_equal(du.P.cpu().numpy(), cu.P) and np.array_equal(du.episode.cpu().numpy(), cu.episode)
    psi = 3
    Qext = rng.rand(tb.S, tb.A * psi).astype(np.float32)

    def dev_planner_p(ag, idx):
        ag.Q[idx.long()] = torch.from_numpy(Qext).cuda()

    def cpu_planner_p(loops, idx):
        loops.Q[idx] = Qext

    dp_ = al.PSRLContinuous(seed, tb, 801, psi_weight=0.015, eta_weight=1e-9, n_loops=N, planner=dev_planner_p, **ex)
    assert dp_._psi == psi
    cp = orc.PSRLCLoops(host_tables(tb), N, psi, seed=seed, planner=cpu_planner_p, **ex)
    assert np.array_equal(torch.cat([dp_.steps(300, trace=True), dp_.steps(500, trace=True)]).cpu().numpy(), cp.steps(800, trace=True))
    assert np.array_equal(dp_.dir_hyper.cpu().numpy(), cp.dir_hyper) and np.array_equal(dp_.episode.cpu().numpy(), cp.episode)
