"""GPU: the model-based continuous agents as device loops (csrc/continuous_agents.cu, csrc/extended_vi.cu through
colosseum_b200.agent_loop) against the oracle -- itself pinned to the reference's agent classes by
tests/golden/make_ucrl2_golden.py / make_psrlc_golden.py -- and, where the reference package is staged, against the
unmodified reference class directly."""
import importlib
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN, load_instance
from colosseum_b200.tables import MDPTables
from oracle import oracle as orc

sys.path.insert(0, GOLDEN)
from make_qlearning_golden import host_tables  # noqa: E402
from make_ucrl2_golden import CASES as UCASES, N_LOOPS as UN, N_STEPS as UT, SEED as USEED  # noqa: E402

pytestmark = pytest.mark.gpu


def _ucrl2_pairs(dev, cpu):
    return [("N", dev.N, cpu.Nsas), ("Nsa", dev.Nsa, cpu.Nsa), ("P", dev.P, cpu.P),
            ("est_r", dev.estimated_rewards, cpu.est_r), ("var_r", dev.variance_proxy_reward, cpu.var_r),
            ("hold", dev.estimated_holding_times, cpu.hold), ("iteration", dev.iteration, cpu.iteration),
            ("episode", dev.episode, cpu.episode), ("delta", dev.delta, cpu.delta), ("state", dev.state, cpu.state),
            ("time", dev.time, cpu.t), ("cum_reward", dev.cumulative_reward, cpu.cum_reward), ("nu", dev.nu, cpu.nu),
            ("ep_len", dev.ep_len, cpu.ep_len)]


@pytest.mark.parametrize("name,inst,kw", UCASES, ids=[c[0] for c in UCASES])
def test_ucrl2_kernels_equal_oracle_bit_for_bit(name, inst, kw):
    """Steps, artificial-episode ends, bounds and model updates of N loops: kernel == oracle bit for bit.  Both sides act
    on the SAME Q after every episode end (the oracle's extended VI, looked up by (loop, episode)), so one flipped
    near-tie cannot hide behind a diverged trajectory; the device planner is checked against that Q at every call."""
    import torch

    import colosseum_b200.agent_loop as al

    tb = MDPTables.from_golden(load_instance(inst))
    N, n_steps, seed = 37, 1500, 11
    cpu = orc.UCRL2Loops(host_tables(tb), N, n_steps + 1, seed=seed, record=True, **kw)
    tr_c = cpu.steps(n_steps, trace=True)
    worst = {"beta": 0.0, "Q": 0.0, "calls": 0}

    def planner(ag, idx, br, bp):
        ids = idx.cpu().numpy()
        ep = ag.episode.cpu().numpy()
        want = [cpu.history[(int(i), int(ep[i]))] for i in ids]
        for k, (_, br_c, bp_c) in enumerate(want):
            worst["beta"] = max(worst["beta"], float(np.abs(br[k].cpu().numpy() / br_c - 1).max()),
                                float(np.abs(bp[k].cpu().numpy() / bp_c - 1).max()))
        # the device planner on the same model, into scratch tables
        Qs, Vs = torch.zeros_like(ag.Q), torch.zeros_like(ag.V)
        span, iters, status = ag.solve_optimistic_model(idx, br, bp, Qs, Vs)
        assert int(status.max()) == 0
        Qd = Qs[idx.long()].cpu().numpy()
        Qc = np.stack([w[0] for w in want])
        worst["Q"] = max(worst["Q"], float(np.abs(Qd - Qc).max()))
        worst["calls"] += len(ids)
        ag.Q[idx.long()] = torch.from_numpy(Qc).cuda()

    dev = al.UCRL2Continuous(seed, tb, n_steps + 1, n_loops=N, planner=planner, **kw)
    tr_d = torch.cat([dev.steps(400, trace=True), dev.steps(n_steps - 400, trace=True)]).cpu().numpy()
    assert np.array_equal(tr_d, tr_c)
    for f, d, c in _ucrl2_pairs(dev, cpu):
        assert np.array_equal(d.cpu().numpy(), c), f
    assert worst["calls"] > 10 * N
    assert worst["beta"] < 1e-14, worst  # log() is the one operation that is not correctly rounded on both sides
    assert worst["Q"] < 2e-3, worst      # the extended VI stops at eps = 1e-3 (infinite_horizon.py:111)


def test_ucrl2_golden_trace_and_reference_tables():
    """the complete device agent (its own batched extended VI as the planner): every loop's trajectory is replayed by
    the oracle acting on the DEVICE's Q of each episode -- same trajectory, same episode ends, bit-identical model tables
    -- i.e. the device loop is a valid run of the agent the golden file pins to the reference class."""
    import torch

    import colosseum_b200.agent_loop as al

    name, inst, kw = UCASES[1]
    tb = MDPTables.from_golden(load_instance(inst))
    N, n_steps, seed = 64, 3000, 5
    hist = {}

    class Recording(al.UCRL2Continuous):
        def episode_end_update(self, idx, update_model=True):
            super().episode_end_update(idx, update_model)
            ep = self.episode.cpu().numpy()
            Q = self.Q[idx.long()].cpu().numpy()
            for k, i in enumerate(idx.cpu().numpy()):
                hist[(int(i), int(ep[i]))] = Q[k]

    dev = Recording(seed, tb, n_steps + 1, n_loops=N, **kw)
    tr_d = dev.steps(n_steps, trace=True).cpu().numpy()
    assert dev.rounds > 20 and dev.evi_iterations > 0

    def planner(i, episode, P, est, br, bp, r_max):
        return 0.0, hist[(i, episode)], np.zeros(tb.S, np.float32)

    cpu = orc.UCRL2Loops(host_tables(tb), N, n_steps + 1, seed=seed, planner=planner, **kw)
    tr_c = cpu.steps(n_steps, trace=True)
    assert np.array_equal(tr_d, tr_c)
    for f, d, c in _ucrl2_pairs(dev, cpu):
        assert np.array_equal(d.cpu().numpy(), c), f


def test_ucrl2_device_loop_through_the_reference_class():
    """where the reference package is staged (baseline/_ref on the GPU box): the device loops' trajectories through the
    UNMODIFIED UCRL2Continuous -- its episode ends and model tables must be the device's, bit for bit, and its numba
    extended VI must land within the stopping tolerance of the device's Q."""
    from oracle.reference_import import reference_available

    if not reference_available():
        pytest.skip("the reference package is not staged (baseline/_ref)")
    import colosseum_b200.agent_loop as al
    from make_qlearning_golden import reference_models
    from make_ucrl2_golden import mdp_spec, replay_reference

    reference_models()
    ucrl2 = importlib.import_module("colosseum.agent.agents.infinite_horizon.ucrl2")
    for name, inst, kw in UCASES[:2]:
        tb = MDPTables.from_golden(load_instance(inst))
        N, n_steps, seed = 4, 2500, USEED
        dev = al.UCRL2Continuous(seed, tb, n_steps + 1, n_loops=N, **kw)
        tr = dev.steps(n_steps, trace=True).cpu().numpy()
        for i in (0, N - 1):
            ag, ends = replay_reference(ucrl2, mdp_spec(tb), tr[:, i], n_steps + 1, kw, seed=seed)
            assert int(dev.episode[i]) == ag.episode and int(dev.iteration[i]) == ag.iteration, name
            assert float(dev.delta[i]) == ag.delta
            for d, r in ((dev.N[i], ag.N), (dev.P[i], ag.P), (dev.estimated_rewards[i], ag.estimated_rewards),
                         (dev.variance_proxy_reward[i], ag.variance_proxy_reward),
                         (dev.estimated_holding_times[i], ag.estimated_holding_times)):
                assert np.array_equal(d.cpu().numpy(), r), name
            assert np.abs(dev.Q[i].cpu().numpy() - ag.Q).max() < 2e-3, name


def test_extended_vi_batched_equals_single_instance():
    """colo_extended_vi_batched_f32 (one CTA per listed instance) == colo_extended_vi_f32 (one launch per iteration),
    bit for bit, on the reference's numba goldens (tests/golden/evi.npz) -- listed out of order, with an unlisted
    instance left untouched."""
    import ctypes as C

    import torch

    from colosseum_b200 import _cabi

    gold = np.load(os.path.join(GOLDEN, "evi.npz"))
    lib = _cabi.lib()
    for nm in range(int(gold["n_cases"])):
        T, est = gold[f"P_{nm}"], gold[f"est_{nm}"]
        br, bp = gold[f"beta_r_{nm}"], gold[f"beta_p_{nm}"]
        r_max = 1.0
        S, A, _ = T.shape
        bp = np.ascontiguousarray(bp[..., 0] if bp.ndim == 3 else bp).reshape(S, A)
        # three instances: the golden model, a perturbed one, and one that is not listed
        Ts = torch.from_numpy(np.stack([T, T, T])).cuda()
        ests = torch.from_numpy(np.stack([est, est * 0.5, est])).cuda().float()
        idx = torch.tensor([1, 0], dtype=torch.int32, device="cuda")
        brd = torch.from_numpy(np.stack([br, br])).cuda().double().contiguous()
        bpd = torch.from_numpy(np.stack([bp * 2.0, bp])).cuda().double().contiguous()
        Q = torch.full((3, S, A), -7.0, dtype=torch.float32, device="cuda")
        V = torch.full((3, S), -7.0, dtype=torch.float32, device="cuda")
        span = torch.zeros(2, dtype=torch.float64, device="cuda")
        iters = torch.zeros(2, dtype=torch.int64, device="cuda")
        status = torch.zeros(2, dtype=torch.int32, device="cuda")
        rc = lib.colo_extended_vi_batched_f32(Ts.data_ptr(), ests.data_ptr(), brd.data_ptr(), bpd.data_ptr(),
                                              idx.data_ptr(), 2, S, A, r_max, 1e-3, int(1e6), Q.data_ptr(), V.data_ptr(),
                                              span.data_ptr(), iters.data_ptr(), status.data_ptr(), _cabi.current_stream())
        _cabi.check(rc, "colo_extended_vi_batched_f32")
        assert status.cpu().tolist() == [0, 0]
        assert float(Q[2].max()) == -7.0 and float(V[2].max()) == -7.0
        work = torch.empty(lib.colo_extended_vi_work_bytes(S, 0), dtype=torch.uint8, device="cuda")
        for k, inst in enumerate((1, 0)):
            Q1 = torch.empty((S, A), dtype=torch.float32, device="cuda")
            V1 = torch.empty(S, dtype=torch.float32, device="cuda")
            out = (C.c_double * 2)()
            rc = lib.colo_extended_vi_f32(Ts[inst].data_ptr(), ests[inst].data_ptr(), brd[k].data_ptr(), bpd[k].data_ptr(),
                                          S, A, r_max, 1e-3, int(1e6), Q1.data_ptr(), V1.data_ptr(), out, work.data_ptr(),
                                          _cabi.current_stream())
            _cabi.check(rc, "colo_extended_vi_f32")
            assert torch.equal(Q1, Q[inst]) and torch.equal(V1, V[inst]), nm
            assert out[0] == float(span[k]) and int(out[1]) == int(iters[k]), nm
        # and the golden of the reference's numba run, within the stopping tolerance
        assert np.abs(Q[0].cpu().numpy() - gold[f"Q_{nm}_loose"]).max() < 2e-3 + 3e-6, nm


def test_ucrl2_learns_river_swim():
    """256 loops on RiverSwimContinuous (optimal average reward 0.889, uniformly random policy 0.017): the batch's
    reward rate over the last fifth of 30,000 steps is above 0.7 (measured 0.85; first fifth 0.61)."""
    import colosseum_b200.agent_loop as al

    tb = MDPTables.from_golden(load_instance("riverswimcontinuous_ergo0"))
    T = 30000
    ag = al.UCRL2Continuous(0, tb, T + 1, alpha_r=0.1, alpha_p=0.05, n_loops=256)
    ag.steps(T * 4 // 5)
    before = float(ag.cumulative_reward.mean())
    ag.steps(T // 5)
    rate = (float(ag.cumulative_reward.mean()) - before) / (T // 5)
    assert rate > 0.7, rate
    assert int(ag.time.min()) == int(ag.time.max()) == T + 1 and int(ag.ended.max()) == 0
