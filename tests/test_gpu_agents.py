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
