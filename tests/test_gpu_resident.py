"""GPU parity of the on-chip resident solver (csrc/resident.cu, `colo_resident_*`): one launch per solve, T kept in
the shared memory of a thread-block cluster.  It must compute exactly what the streaming path computes sweep by
sweep (same synchronous recurrence, infinite_horizon.py:121-184 / finite_horizon.py:11-42 / diameter.py:76-106);
only the summation order inside a row differs, hence tolerances at rounding level rather than bit equality."""
import ctypes as C

import numpy as np
import pytest

from conftest import load_instance
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
GAM = float(np.float32(0.99))  # the reference casts gamma to float32 (infinite_horizon.py:127)


@pytest.fixture(scope="module")
def env():
    import torch

    import colosseum_b200.dynamic_programming as dp
    from colosseum_b200 import _cabi

    return torch, dp, _cabi


def synth(seed, S, A, alpha=0.3):
    rs = np.random.RandomState(seed)
    T = rs.dirichlet(np.ones(S) * alpha, size=(S, A)).astype(np.float32)
    R = rs.uniform(0, 1, size=(S, A)).astype(np.float32)
    return T, R


def resident(env, T, R, *, n_iter, eps=0.0, NV=1, f64=False, pins=None, fold=0, gamma=0.99, r_const=0.0, pi=None,
             episodic_H=0, max_abs=0.0, overflow_signed=0, want_q=True):
    torch, dp, _cabi = env
    lib = _cabi.lib()
    Td = torch.from_numpy(np.ascontiguousarray(T, np.float32)).cuda()
    batched = Td.dim() == 4
    S, A = Td.shape[-3], Td.shape[-2]
    B = (len(pins) // 4) if NV == 4 else (Td.shape[0] if batched else 1)
    Rd = None if R is None else torch.from_numpy(np.ascontiguousarray(R, np.float32)).cuda()
    pid = None if pi is None else torch.from_numpy(np.ascontiguousarray(pi, np.float32)).cuda()
    pind = None if pins is None else torch.from_numpy(np.ascontiguousarray(pins, np.int32)).cuda()
    vd = torch.float64 if f64 else torch.float32
    L = episodic_H + 1 if episodic_H else 1
    V = torch.zeros((B * NV, L, S), dtype=vd, device="cuda")
    Q = torch.zeros((B * NV, L, S, A), dtype=vd, device="cuda") if want_q else None
    status = torch.full((B,), -7, dtype=torch.int32, device="cuda")
    iters = torch.zeros(B * NV, dtype=torch.int64, device="cuda")
    a = _cabi.ResidentArgs()
    a.T, a.R, a.pi, a.V, a.Q = _cabi.ptr(Td), _cabi.ptr(Rd), _cabi.ptr(pid), _cabi.ptr(V), _cabi.ptr(Q)
    a.t_stride = 0 if NV == 4 else S * A * S
    a.r_stride = S * A
    a.B, a.S, a.A, a.NV, a.fold = B, S, A, NV, fold
    a.gamma, a.r_const, a.eps, a.max_iter = gamma, r_const, eps, n_iter
    a.max_abs, a.overflow_signed, a.episodic_H = max_abs, overflow_signed, episodic_H
    a.pin_index = _cabi.ptr(pind)
    a.iters_out, a.status_out = _cabi.ptr(iters), _cabi.ptr(status)
    fn = lib.colo_resident_solve_f64acc if f64 else lib.colo_resident_solve_f32
    rc = fn(C.byref(a), _cabi.current_stream())
    assert rc == 0, _cabi.last_error()
    torch.cuda.synchronize()
    return (V.cpu().numpy().squeeze(1) if L == 1 else V.cpu().numpy(),
            None if Q is None else (Q.cpu().numpy().squeeze(1) if L == 1 else Q.cpu().numpy()),
            status.cpu().numpy(), iters.cpu().numpy())


def fits(env, S, A, NV=1, f64=False):
    _, _, _cabi = env
    cs = C.c_int(0)
    ok = _cabi.lib().colo_resident_fits(S, A, NV, int(f64), C.byref(cs))
    return ok, cs.value


def test_fits_plan(env):
    assert fits(env, 16, 5) == (1, 1)
    ok, c = fits(env, 465, 2)  # C2's MDP: 1.73 MB of T -> a 16-CTA cluster
    assert ok == 1 and c == 16
    assert fits(env, 512, 4)[0] == 0  # C4's shape (4.19 MB) streams from HBM instead
    assert fits(env, 40000, 8)[0] == 0


# cluster sizes 1 .. 16, S % 4 != 0, A > 4 (two register tiles), A = 1
@pytest.mark.parametrize("S,A", [(16, 5), (23, 4), (61, 3), (108, 6), (210, 2), (333, 1), (465, 2), (400, 3)])
@pytest.mark.parametrize("f64", [False, True])
def test_fixed_sweeps_match_streaming_path(env, S, A, f64):
    """n sweeps inside one resident launch == n launches of the streaming backup kernel"""
    torch, dp, _ = env
    assert fits(env, S, A, 1, f64)[0] == 1
    T, R = synth(S * 31 + A, S, A)
    n = 37
    V, Q, status, iters = resident(env, T, R, n_iter=n, f64=f64, gamma=GAM)
    assert status[0] == 2 and iters[0] == n  # eps = 0 never converges: COLO_MAX_ITER after exactly n sweeps
    vi = dp.BatchedValueIteration(T, R, gamma=0.99, precision="f64" if f64 else "f32")
    vi.sweep(n)
    tol = 1e-12 if f64 else 2e-6
    np.testing.assert_allclose(V[0], vi.values[0].cpu().numpy(), rtol=tol, atol=tol)
    np.testing.assert_allclose(Q[0], vi.Q[0].cpu().numpy(), rtol=tol, atol=tol)
    Qo, Vo = orc.jacobi_sweeps_f64(T, R, np.zeros(S), n, gamma=GAM)
    tol = 1e-9 if f64 else 1e-4
    np.testing.assert_allclose(V[0], Vo, rtol=tol)
    np.testing.assert_allclose(Q[0], Qo, rtol=tol)


@pytest.mark.parametrize("f64", [False, True])
def test_batched_instances_stop_independently(env, f64):
    """cluster b iterates instance b to its own stopping sweep (infinite_horizon.py:140-141)"""
    B, S, A = 9, 40, 3
    Ts, Rs = zip(*[synth(b, S, A) for b in range(B)])
    T, R = np.stack(Ts), np.stack(Rs)
    R[4] *= 1e-3
    eps = 1e-9 if f64 else 1e-5
    V, Q, status, iters = resident(env, T, R, n_iter=10**6, eps=eps, gamma=0.9, f64=f64)
    assert (status == 0).all() and iters[4] < iters.max()
    for b in range(B):
        Qo, Vo, n_o = orc.discounted_f64(T[b], R[b], gamma=0.9, tol=eps)
        tol = 1e-6 if f64 else 1e-4
        np.testing.assert_allclose(V[b], Vo, rtol=tol, atol=2 * eps)
        np.testing.assert_allclose(Q[b], Qo, rtol=tol, atol=2 * eps)
        assert abs(int(iters[b]) - int(n_o)) <= 1


@pytest.mark.parametrize("name", ["doc_simplegrid4", "taxicontinuous_ergo0", "deepsea20_prand"])
def test_hitting_time_tiles(env, name):
    """NV = 4: four targets per cluster share every T quad (diameter.py:76-106), vs the per-target oracle"""
    T = load_instance(name)["T"]
    S = T.shape[0]
    K4 = (S + 3) // 4 * 4
    pins = np.minimum(np.arange(K4), S - 1).astype(np.int32)
    V, _, status, iters = resident(env, T, None, n_iter=10**6, eps=1e-10, NV=4, f64=True, pins=pins, fold=2,
                                   gamma=1.0, r_const=1.0, want_q=False)
    assert (status == 0).all()
    d, E, _ = orc.diameter_continuous_f64(T, return_E=True)  # E[k,s]: optimal expected hitting time of k from s
    np.testing.assert_allclose(V[:S], E, rtol=1e-6, atol=1e-8)
    assert abs(V[:S].max() - d) < 1e-6 * d


@pytest.mark.parametrize("f64", [False, True])
def test_episodic_layers_and_policy(env, f64):
    """episodic_H: exactly H sweeps, layer H-1-i stored by sweep i (finite_horizon.py:11-42), max and pi folds"""
    S, A, H = 50, 4, 9
    T, R = synth(5, S, A)
    pi = np.random.RandomState(2).dirichlet(np.ones(A), size=(H, S)).astype(np.float32)
    tol = 1e-9 if f64 else 2e-5
    V, Q, status, _ = resident(env, T, R, n_iter=H, gamma=1.0, episodic_H=H, f64=f64)
    Qo, Vo = orc.episodic_f64(H, T, R)
    assert status[0] == 0
    np.testing.assert_allclose(V[0][:H], Vo[:H], rtol=tol)
    np.testing.assert_allclose(Q[0][:H], Qo[:H], rtol=tol)
    assert (V[0][H] == 0).all()  # layer H is the caller's (zero) terminal layer
    V, Q, status, _ = resident(env, T, R, n_iter=H, gamma=1.0, episodic_H=H, fold=1, pi=pi, f64=f64)
    Qo, Vo = orc.episodic_f64(H, T, R, pi=pi)
    np.testing.assert_allclose(V[0][:H], Vo[:H], rtol=tol)
    np.testing.assert_allclose(Q[0][:H], Qo[:H], rtol=tol)


def test_overflow_status(env):
    T, R = synth(3, 30, 2)
    _, _, status, iters = resident(env, T, R, n_iter=10**6, eps=1e-6, max_abs=5.0)
    assert status[0] == 1 and iters[0] < 50  # |V| passes 5 after a handful of sweeps (infinite_horizon.py:136-138)
    _, _, status, _ = resident(env, T, -R, n_iter=7, gamma=1.0, episodic_H=7, max_abs=1.5, overflow_signed=1)
    assert status[0] == 0  # signed test (finite_horizon.py:24-25): V <= 0 never exceeds max_value
    _, _, status, _ = resident(env, T, R, n_iter=7, gamma=1.0, episodic_H=7, max_abs=1.5, overflow_signed=1)
    assert status[0] == 1
