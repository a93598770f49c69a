"""GPU parity: the CUDA Bellman backups (through the C ABI) against the CPU oracle and the reference goldens.

Tolerances (BASELINE.json north_star): value functions within 1e-6 relative in fp64 mode, 1e-4 in fp32 mode, at
the fixed point of the recurrence (the reference's dense kernel sweeps in place, the GPU sweeps synchronously;
DESIGN.md explains why parity is defined at the fixed point)."""
import numpy as np
import pytest

from conftest import CONTINUOUS, EPISODIC, load_instance
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

RTOL64, RTOL32 = 1e-6, 1e-4
GAM = float(np.float32(0.99))


@pytest.fixture(scope="module")
def dp():
    import colosseum_b200.dynamic_programming as dp

    return dp


def synth(seed, S, A, alpha=0.05):
    rs = np.random.RandomState(seed)
    T = rs.dirichlet(np.ones(S) * alpha, size=(S, A)).astype(np.float32)
    T = (T / T.sum(-1, keepdims=True, dtype=np.float32)).astype(np.float32)
    R = rs.uniform(0, 1, size=(S, A)).astype(np.float32)
    return T, R


@pytest.mark.parametrize("S,A", [(24, 4), (33, 3), (64, 2), (512, 4), (130, 5), (1, 1), (4, 9)])
@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_single_backup_matches_oracle(dp, S, A, precision):
    """one synchronous sweep (infinite_horizon.py:131-135), vector and scalar load paths (S % 4 != 0), fold max/pi/min"""
    T, R = synth(S * 7 + A, S, A, alpha=0.5)
    rs = np.random.RandomState(1)
    V0 = rs.uniform(-3, 5, S)
    pi = rs.dirichlet(np.ones(A), size=S).astype(np.float32)
    rtol = RTOL64 if precision == "f64" else RTOL32
    for fold, p in ((orc.FOLD_MAX, None), (orc.FOLD_PI, pi), (orc.FOLD_MIN, None)):
        Qo, Vo = orc.jacobi_sweeps_f64(T, R, V0, 1, gamma=0.9, pi=p, fold=fold)
        Q, V, res = dp.bellman_backup(T, R, V0, gamma=0.9, pi=p, fold=fold, precision=precision, return_residual=True)
        np.testing.assert_allclose(Q, Qo, rtol=rtol, atol=rtol)
        np.testing.assert_allclose(V, Vo, rtol=rtol, atol=rtol)
        v0 = V0 if precision == "f64" else V0.astype(np.float32)
        np.testing.assert_allclose(res, np.abs(Vo - v0).max(), rtol=1e-3, atol=1e-5)


@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_discounted_vi_pe_golden_synth(dp, dp_synth, precision):
    g = dp_synth
    rtol = RTOL64 if precision == "f64" else RTOL32
    eps = 1e-10 if precision == "f64" else 1e-5
    for b in range(3):
        T, R, pi = g[f"T_{b}"], g[f"R_{b}"], g[f"pi_{b}"]
        Qo, Vo, _ = orc.discounted_f64(T, R, gamma=GAM, tol=1e-13)
        Q, V = dp.discounted_value_iteration(T, R, 0.99, eps, precision=precision)
        assert Q.shape == Qo.shape and V.shape == Vo.shape
        assert Q.dtype == (np.float64 if precision == "f64" else np.float32)
        np.testing.assert_allclose(V, Vo, rtol=rtol)
        np.testing.assert_allclose(Q, Qo, rtol=rtol)
        # ... and against the reference itself run to eps=1e-6 (its own fp32 noise bounds this one)
        np.testing.assert_allclose(V, g[f"Vt_{b}"], rtol=1e-5, atol=2e-4)
        Qpo, Vpo, _ = orc.discounted_f64(T, R, pi=pi, gamma=GAM, tol=1e-13)
        Qp, Vp = dp.discounted_policy_evaluation(T, R, pi, 0.99, eps, precision=precision)
        np.testing.assert_allclose(Vp, Vpo, rtol=rtol)
        np.testing.assert_allclose(Qp, Qpo, rtol=rtol)
        np.testing.assert_allclose(Vp, g[f"Vp_{b}"], rtol=3e-5)


def test_reference_default_epsilon_semantics(dp, dp_synth):
    """with the reference's default eps=1e-3 the GPU (Jacobi) result is within eps*gamma/(1-gamma) of the fixed
    point, exactly like the reference's own (Gauss-Seidel) result -- both are early-stopped iterates"""
    g = dp_synth
    T, R = g["T_0"], g["R_0"]
    Q, V = dp.discounted_value_iteration(T, R)
    _, Vstar, _ = orc.discounted_f64(T, R, gamma=GAM, tol=1e-13)
    bound = 1e-3 * GAM / (1 - GAM)
    assert np.abs(V - Vstar).max() < bound and np.abs(g["V_0"] - Vstar).max() < bound
    assert V.dtype == np.float32 and Q.dtype == np.float32


def test_overflow_and_max_iter_contracts(dp, dp_synth):
    g = dp_synth
    T, R = g["T_0"], g["R_0"]
    assert dp.discounted_value_iteration(T, R, max_abs_value=5.0) is None  # infinite_horizon.py:136-138
    assert dp.episodic_value_iteration(7, T, R, max_value=1.5) is None  # finite_horizon.py:24-25
    assert dp.episodic_value_iteration(7, T, R, max_value=1e9) is not None
    with pytest.raises(dp.DynamicProgrammingMaxIterationExceeded):  # infinite_horizon.py:142
        dp._solve_discounted(T, R, None, 0.99, 1e-12, None, "f64", max_iter=5)


@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_batched_instances_converge_independently(dp, precision):
    """config C4's shape at test size: a batch of independent MDPs, each stopping at its own sweep"""
    B, S, A = 6, 48, 4
    Ts, Rs = zip(*[synth(b, S, A) for b in range(B)])
    T, R = np.stack(Ts), np.stack(Rs)
    R[3] *= 0.01  # converges much earlier than the others
    eps = 1e-9 if precision == "f64" else 1e-5  # early-stop error <= eps*gamma/(1-gamma)
    Q, V = dp.discounted_value_iteration(T, R, 0.95, eps, precision=precision)
    iters = dp.last_iterations()
    assert len(iters) == B and iters[3] < max(iters)
    rtol = RTOL64 if precision == "f64" else RTOL32
    for b in range(B):
        # same algorithm (synchronous sweeps) and same stopping rule in the oracle: the two runs stop at the same
        # sweep or one apart, i.e. within eps of each other on top of the arithmetic tolerance
        Qo, Vo, _ = orc.discounted_f64(T[b], R[b], gamma=float(np.float32(0.95)), tol=eps)
        np.testing.assert_allclose(V[b], Vo, rtol=rtol, atol=2 * eps)
        np.testing.assert_allclose(Q[b], Qo, rtol=rtol, atol=2 * eps)
        # and each instance is bit-identical to solving it alone (no cross-instance coupling)
        Q1, V1 = dp.discounted_value_iteration(T[b], R[b], 0.95, eps, precision=precision)
        assert np.array_equal(V1, V[b]) and np.array_equal(Q1, Q[b])


@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_episodic_golden(dp, dp_synth, precision):
    g = dp_synth
    rtol = RTOL64 if precision == "f64" else 2e-5
    for b in range(3):
        T, R, H = g[f"T_{b}"], g[f"R_{b}"], int(g[f"H_{b}"])
        Qo, Vo = orc.episodic_f64(H, T, R)
        Q, V = dp.episodic_value_iteration(H, T, R, precision=precision)
        assert Q.shape == (H + 1,) + R.shape and V.shape == (H + 1, R.shape[0])
        assert (V[H] == 0).all() and (Q[H] == 0).all()  # finite_horizon.py:17-18
        np.testing.assert_allclose(V, Vo, rtol=rtol, atol=1e-7)
        np.testing.assert_allclose(Q, Qo, rtol=rtol, atol=1e-7)
        np.testing.assert_allclose(V, g[f"Ve_{b}"], rtol=1e-5, atol=1e-6)  # the reference's own output
        pol = g[f"pol_{b}"]
        Qpo, Vpo = orc.episodic_f64(H, T, R, pi=pol)
        Qp, Vp = dp.episodic_policy_evaluation(H, T, R, pol, precision=precision)
        np.testing.assert_allclose(Vp, Vpo, rtol=rtol, atol=1e-7)
        np.testing.assert_allclose(Vp, g[f"Vpe_{b}"], rtol=1e-5, atol=1e-6)


def test_c1_anchor(dp):
    """BASELINE.json configs[0]: RiverSwimEpisodic size 5, V[0] from SURVEY.md section 8d"""
    g = load_instance("c1_riverswim_epi")
    Q, V = dp.episodic_value_iteration(int(g["H"]), g["T"], g["R"])
    np.testing.assert_allclose(V[0], [0.45454547, 0.36414355, 0.2737823, 0.3346822, 0.4166667], rtol=2e-6)
    np.testing.assert_allclose(V, g["vi_V"], rtol=2e-6, atol=1e-7)
    np.testing.assert_allclose(Q, g["vi_Q"], rtol=2e-6, atol=1e-7)


@pytest.mark.parametrize("name", CONTINUOUS)
def test_continuous_benchmark_instances(dp, name):
    g = load_instance(name)
    T, R = g["T"], g["R"]
    S, A = R.shape
    Qo, Vo, _ = orc.discounted_f64(T, R, gamma=GAM, tol=1e-13)
    Q, V = dp.discounted_value_iteration(T, R, 0.99, 1e-10, precision="f64")
    np.testing.assert_allclose(V, Vo, rtol=RTOL64)
    np.testing.assert_allclose(Q, Qo, rtol=RTOL64, atol=1e-9)
    np.testing.assert_allclose(V, g["vi_tight_V"], rtol=1e-5, atol=3e-4)  # reference at eps=1e-6
    # fp32 mode: same synchronous algorithm and stopping rule as the oracle run below (parity)
    Q32, V32 = dp.discounted_value_iteration(T, R, 0.99, 2e-5, precision="f32")
    _, Vo32, _ = orc.discounted_f64(T, R, gamma=GAM, tol=2e-5)
    np.testing.assert_allclose(V32, Vo32, rtol=RTOL32, atol=4e-5)
    assert np.abs(V32 - Vo).max() < 2e-5 * GAM / (1 - GAM) + RTOL32 * np.abs(Vo).max()  # early-stop bound
    pi = np.ones((S, A), np.float32) / A
    _, Vpo, _ = orc.discounted_f64(T, R, pi=pi, gamma=GAM, tol=1e-13)
    _, Vp = dp.discounted_policy_evaluation(T, R, pi, precision="f64", epsilon=1e-10)
    np.testing.assert_allclose(Vp, Vpo, rtol=RTOL64, atol=1e-9)
    np.testing.assert_allclose(Vp, g["pe_V"], rtol=5e-5, atol=1e-5)  # the reference's own output


@pytest.mark.parametrize("name", EPISODIC)
def test_episodic_benchmark_instances(dp, name):
    g = load_instance(name)
    T, R, H = g["T"], g["R"], int(g["H"])
    S, A = R.shape
    Q, V = dp.episodic_value_iteration(H, T, R, precision="f64")
    Qo, Vo = orc.episodic_f64(H, T, R)
    np.testing.assert_allclose(V, Vo, rtol=RTOL64, atol=1e-9)
    np.testing.assert_allclose(Q, Qo, rtol=RTOL64, atol=1e-9)
    np.testing.assert_allclose(V, g["vi_V"], rtol=2e-5, atol=1e-6)
    pol = np.ones((H, S, A), np.float32) / A
    _, Vp = dp.episodic_policy_evaluation(H, T, R, pol, precision="f32")
    np.testing.assert_allclose(Vp, g["pe_V"], rtol=2e-5, atol=1e-6)


def test_full_size_c4_properties(dp):
    """BASELINE.json configs[3] shape (S=512, A=4), a slice of the batch at full per-MDP size: size-independent
    properties -- Bellman residual of the returned V is below eps, V is monotone in R, and a converged batch
    entry equals the stand-alone solve."""
    import torch

    B, S, A = 8, 512, 4
    gen = torch.Generator(device="cuda").manual_seed(0)
    T = torch._standard_gamma(torch.full((B, S, A, S), 0.05, device="cuda"), generator=gen).float() + 1e-30
    T = (T / T.sum(-1, keepdim=True)).contiguous()
    R = torch.rand((B, S, A), device="cuda", generator=gen)
    Q, V = dp.discounted_value_iteration(T, R, 0.99, 1e-4, precision="f32")
    assert V.is_cuda and V.shape == (B, S)
    Q2, V2, res = dp.bellman_backup(T, R, V, gamma=GAM, precision="f32", return_residual=True)
    assert float(res.max()) < 1e-4 * 1.5  # one more sweep moves V by less than eps
    Qhi, Vhi = dp.discounted_value_iteration(T, R + 0.1, 0.99, 1e-4, precision="f32")
    assert bool((Vhi >= V).all())  # monotone in R
    np.testing.assert_allclose((Vhi - V).cpu().numpy(), 0.1 / (1 - GAM), rtol=2e-3)  # V(R+c) = V(R) + c/(1-gamma)
    # oracle on one full-size instance
    _, Vo, _ = orc.discounted_f64(T[5].cpu().numpy(), R[5].cpu().numpy(), gamma=GAM, tol=1e-4)
    np.testing.assert_allclose(V[5].cpu().numpy(), Vo, rtol=RTOL32, atol=2e-4)


def test_get_policy_from_q_values(dp):
    Q = np.array([[0.0, 1.0, 0.5], [2.0, 2.0, 1.0], [3.0, 1.0, 3.0]], np.float32)
    pi = dp.get_policy_from_q_values(Q)
    assert pi.dtype == np.int32 and pi[0] == 1 and pi[1] in (0, 1) and pi[2] in (0, 2)
    X = dp.get_policy_from_q_values(Q, True)
    assert X.shape == Q.shape and (X.sum(-1) == 1).all() and (Q[np.arange(3), X.argmax(-1)] == Q.max(-1)).all()
    X3 = dp.get_policy_from_q_values(np.stack([Q, Q]), True)
    assert X3.shape == (2, 3, 3)


@pytest.mark.parametrize("precision", ["f32", "f64"])
def test_extended_value_iteration(dp, precision):
    """UCRL2's optimistic VI (infinite_horizon.py:67-118 + _max_proba) vs the reference's own numba run (goldens) and
    the oracle; same synchronous iterates, so the bar is the stopping tolerance itself"""
    import os

    from conftest import GOLDEN

    g = np.load(os.path.join(GOLDEN, "evi.npz"))
    for i in range(int(g["n_cases"])):
        args = (g[f"P_{i}"], g[f"est_{i}"], g[f"beta_r_{i}"], g[f"beta_p_{i}"], 1.0)
        for tag, eps in (("loose", 1e-3), ("tight", 1e-5)):
            span, Q, V, it = dp.extended_value_iteration(*args, eps, precision=precision, return_iterations=True)
            so, Qo, Vo, ito = orc.extended_vi_f32(*args, eps)
            assert Q.dtype == (np.float64 if precision == "f64" else np.float32) and Q.shape == Qo.shape
            assert abs(it - ito) <= 1 + ito // 50
            for ref_span, ref_Q, ref_V in ((so, Qo, Vo), (float(g[f"span_{i}_{tag}"]), g[f"Q_{i}_{tag}"], g[f"V_{i}_{tag}"])):
                assert abs(span - ref_span) < 2 * eps + 1e-5
                np.testing.assert_allclose(Q, ref_Q, atol=2 * eps + 3e-6)
                np.testing.assert_allclose(V, ref_V, atol=2 * eps + 3e-6)
    # [S,A,S]-shaped beta_p: only element 0 of the last axis is read (:230); iteration cap -> None
    bp3 = np.repeat(g["beta_p_0"], g["P_0"].shape[0], axis=2) * np.linspace(1, 2, g["P_0"].shape[0])
    r1 = dp.extended_value_iteration(g["P_0"], g["est_0"], g["beta_r_0"], bp3, 1.0, 1e-3, precision=precision)
    r2 = dp.extended_value_iteration(g["P_0"], g["est_0"], g["beta_r_0"], g["beta_p_0"], 1.0, 1e-3, precision=precision)
    assert r1[0] == r2[0] and np.array_equal(r1[1], r2[1])
    assert dp.extended_value_iteration(g["P_0"], g["est_0"], g["beta_r_0"], g["beta_p_0"], 1.0, 1e-9, precision=precision,
                                       max_iter=3) is None


def test_gauss_seidel_reproduces_the_reference_iterate(dp, dp_synth):
    """sweep_order='gauss_seidel': the reference's OWN in-place iterate (infinite_horizon.py:121-142, :167-184), i.e. the
    numbers its numba kernel returns at its default stopping sweep -- against the goldens recorded from the reference
    and the loop-for-loop oracle.  The only difference left is the summation order of a row dot product, which can
    move the stopping sweep by one: the bar is the stopping tolerance itself."""
    g = dp_synth
    for b in range(3):
        T, R, pi = g[f"T_{b}"], g[f"R_{b}"], g[f"pi_{b}"]
        Q, V = dp.discounted_value_iteration(T, R, sweep_order="gauss_seidel")  # reference defaults: 0.99, 1e-3
        it = dp.last_iterations()[0]
        Qo, Vo, ito = orc.discounted_gs_f32(T, R, gamma=0.99, eps=1e-3)
        assert V.dtype == np.float32 and abs(it - ito) <= 1
        for ref_Q, ref_V in ((Qo, Vo), (g[f"Q_{b}"], g[f"V_{b}"])):
            np.testing.assert_allclose(V, ref_V, atol=1.5e-3)
            np.testing.assert_allclose(Q, ref_Q, atol=1.5e-3)
        # ... while the synchronous sweep, stopped by the same rule, sits elsewhere (why the mode exists)
        _, Vj = dp.discounted_value_iteration(T, R)
        assert np.abs(Vj - g[f"V_{b}"]).max() > np.abs(V - g[f"V_{b}"]).max()
        # tight epsilon: both orders meet at the fixed point
        Qt, Vt = dp.discounted_value_iteration(T, R, 0.99, 1e-6, sweep_order="gauss_seidel")
        np.testing.assert_allclose(Vt, g[f"Vt_{b}"], rtol=2e-5, atol=1e-4)
        Qp, Vp = dp.discounted_policy_evaluation(T, R, pi, sweep_order="gauss_seidel", precision="f64")
        np.testing.assert_allclose(Vp, g[f"Vp_{b}"], rtol=3e-5)
    # reference goldens of real MDP instances (their vi_Q / vi_V are the reference's default-epsilon outputs)
    for name in ("doc_simplegrid4", "taxicontinuous_ergo0", "deepsea20_prand"):
        gi = load_instance(name)
        Q, V = dp.discounted_value_iteration(gi["T"], gi["R"], sweep_order="gauss_seidel")
        np.testing.assert_allclose(V, gi["vi_V"], atol=1.5e-3)
        np.testing.assert_allclose(Q, gi["vi_Q"], atol=1.5e-3)
    # contracts: overflow -> None (infinite_horizon.py:136-138); batches; odd S (scalar loads)
    assert dp.discounted_value_iteration(g["T_0"], g["R_0"], max_abs_value=5.0, sweep_order="gauss_seidel") is None
    Tb = np.stack([g["T_1"]] * 5); Rb = np.stack([g["R_1"] * (1 + 0.1 * i) for i in range(5)]).astype(np.float32)
    Qb, Vb = dp.discounted_value_iteration(Tb, Rb, sweep_order="gauss_seidel")
    for i in range(5):
        _, Vi, _ = orc.discounted_gs_f32(Tb[i], Rb[i], gamma=0.99, eps=1e-3)
        np.testing.assert_allclose(Vb[i], Vi, atol=1.5e-3)
    with pytest.raises(dp.DynamicProgrammingMaxIterationExceeded):
        dp._solve_discounted(g["T_0"], g["R_0"], None, 0.99, 1e-12, None, "f64", max_iter=5, sweep_order="gauss_seidel")


@pytest.mark.parametrize("S,A,H", [(5, 2, 5), (36, 2, 8), (108, 6, 7), (130, 3, 4), (200, 4, 3), (465, 2, 3)])
@pytest.mark.parametrize("precision", ["f32", "f64"])
def test_batched_small_episodic_solves(dp, S, A, H, precision):
    """B >= 16 instances take the one-CTA-per-MDP kernel (all layers in one launch; 4/8/16/32 lanes per row, 128-bit
    or scalar loads): VI and PE against the fp64 oracle per instance, and against the per-instance (B = 1) path."""
    import torch

    B = 19
    rs = np.random.RandomState(S)
    T = rs.dirichlet(np.ones(S) * 0.3, size=(B, S, A)).astype(np.float32)
    R = rs.rand(B, S, A).astype(np.float32)
    pol = rs.dirichlet(np.ones(A), size=(B, H, S)).astype(np.float32)
    Td, Rd, pd = torch.from_numpy(T).cuda(), torch.from_numpy(R).cuda(), torch.from_numpy(pol).cuda()
    Q, V = dp.episodic_value_iteration(H, Td, Rd, precision=precision)
    Qp, Vp = dp.episodic_policy_evaluation(H, Td, Rd, pd, precision=precision)
    assert tuple(Q.shape) == (B, H + 1, S, A) and tuple(V.shape) == (B, H + 1, S)
    tol = 1e-9 if precision == "f64" else 2e-5
    for b in (0, 7, B - 1):
        Qo, Vo = orc.episodic_f64(H, T[b], R[b])
        np.testing.assert_allclose(Q[b].cpu().numpy(), Qo, rtol=tol, atol=tol)
        np.testing.assert_allclose(V[b].cpu().numpy(), Vo, rtol=tol, atol=tol)
        Qpo, Vpo = orc.episodic_f64(H, T[b], R[b], pi=pol[b])
        np.testing.assert_allclose(Vp[b].cpu().numpy(), Vpo, rtol=tol, atol=tol)
        Q1, V1 = dp.episodic_value_iteration(H, Td[b], Rd[b], precision=precision)
        np.testing.assert_allclose(V[b].cpu().numpy(), V1.cpu().numpy(), rtol=tol, atol=tol)
    assert float(Q[:, H].abs().max()) == 0.0 and float(V[:, H].abs().max()) == 0.0


def test_tma_staged_backup_matches_the_streaming_kernel(tmp_path):
    """The TMA-staged variant of the synchronous sweep (COLO_BACKUP_TMA=1: cp.async.bulk ring + mbarriers, DESIGN 4.1;
    the switch is read once per process, hence the child processes) returns the values of the shipped LDG kernel --
    fixed sweeps of max / policy-weighted backups with per-instance freezing, Q stored -- and really ran."""
    import os
    import subprocess
    import sys

    child = r"""
import sys, numpy as np, torch
sys.path.insert(0, sys.argv[1])
import colosseum_b200.dynamic_programming as dp
from colosseum_b200 import _cabi
B, S, A = 80, 256, 4
gen = torch.Generator(device="cuda").manual_seed(7)
T = torch._standard_gamma(torch.full((B, S, A, S), 0.05, device="cuda"), generator=gen).float() + 1e-30
T = (T / T.sum(-1, keepdim=True)).contiguous()
R = torch.rand((B, S, A), device="cuda", generator=gen)
vi = dp.BatchedValueIteration(T, R, gamma=0.99, precision="f32")
vi.sweep(25)
Q, V = dp.discounted_value_iteration(T, R, 0.99, 1e-3, precision="f32")
pi = torch.softmax(R * 3, -1).contiguous()
Qp, Vp = dp.discounted_policy_evaluation(T, R, pi, 0.99, 1e-5, precision="f32")
torch.cuda.synchronize()
np.savez(sys.argv[2], V25=vi.values.cpu().numpy(), Q=Q.cpu().numpy(), V=V.cpu().numpy(), Vp=Vp.cpu().numpy(),
         tma=np.int64(_cabi.lib().colo_backup_tma_sweeps()))
"""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = {}
    for tag, env in (("ldg", {}), ("tma", {"COLO_BACKUP_TMA": "1"})):
        path = str(tmp_path / f"{tag}.npz")
        e = {k: v for k, v in os.environ.items() if not k.startswith("COLO_BACKUP_TMA")}
        r = subprocess.run([sys.executable, "-c", child, root, path], env={**e, **env}, capture_output=True, text=True,
                           timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        outs[tag] = np.load(path)
    assert int(outs["ldg"]["tma"]) == 0 and int(outs["tma"]["tma"]) >= 25
    for k in ("V25", "Q", "V", "Vp"):
        np.testing.assert_allclose(outs["tma"][k], outs["ldg"][k], rtol=2e-6, atol=1e-6, err_msg=k)
