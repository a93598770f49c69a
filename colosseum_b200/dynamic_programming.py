"""B200 twin of `colosseum.dynamic_programming` (same names, argument meaning and error behaviour).

Reference surface mirrored here (all paths under /root/reference/colosseum/):
  dynamic_programming/__init__.py:5-14      DP_MAX_ITERATION, the four public entry points
  dynamic_programming/finite_horizon.py:11-42   episodic_value_iteration / episodic_policy_evaluation
  dynamic_programming/infinite_horizon.py:14-64,121-184,208-219   discounted VI / PE / PI
  dynamic_programming/utils.py:8,75-100     DynamicProgrammingMaxIterationExceeded, get_policy_from_q_values

Every function accepts what the reference accepts (C-contiguous float32 numpy arrays, never mutated) and returns
freshly allocated numpy arrays of the reference's shapes; it also accepts CUDA torch tensors (kept resident,
results returned as CUDA tensors) and a leading batch dimension of independent MDP instances.  All arithmetic
runs in the hand-written sm_100a kernels behind the C ABI (include/colosseum_b200.h); there is no CPU path.

Iteration order: the reference's dense kernel sweeps states in place (Gauss-Seidel); the GPU sweeps are
synchronous (Jacobi).  Both converge to the same fixed point, which is where parity is defined; the
`epsilon` stopping rule (max|dV| < epsilon after a sweep) is the reference's.
"""
import ctypes as C

import numpy as np

from . import _cabi
from ._cabi import FOLD_MAX, FOLD_MIN, FOLD_PI  # noqa: F401

DP_MAX_ITERATION = int(1e6)
ARGMAX_SEED = 42

_PRECISION = "f32"  # the reference's arithmetic type; "f64" = fp64 accumulation / fp64 V,Q (1e-6 parity mode)
_SWEEP_ORDER = "jacobi"  # "jacobi": synchronous sweeps (fast; fixed-point parity); "gauss_seidel": the reference's own
#                          in-place iterate (its early-stopped numbers, one warp per instance)


class DynamicProgrammingMaxIterationExceeded(Exception):
    """colosseum/dynamic_programming/utils.py:8"""


def set_precision(p):
    """'f32' (reference arithmetic type) or 'f64' (fp64 accumulation; V, Q returned as float64)."""
    global _PRECISION
    assert p in ("f32", "f64")
    _PRECISION = p


def get_precision():
    return _PRECISION


def set_sweep_order(order):
    """'jacobi' (default) or 'gauss_seidel' (the reference's in-place sweeps, infinite_horizon.py:131-135: the
    discounted solvers then return the reference's early-stopped iterates instead of synchronous ones)."""
    global _SWEEP_ORDER
    assert order in ("jacobi", "gauss_seidel")
    _SWEEP_ORDER = order


def get_sweep_order():
    return _SWEEP_ORDER


def _torch():
    import torch

    return torch


def _is_tensor(x):
    return type(x).__module__.startswith("torch")


def to_device(x, np_dtype=np.float32):
    """numpy (caller-owned, not mutated) -> CUDA tensor; CUDA tensors pass through (made contiguous)."""
    torch = _torch()
    _cabi.require_cuda()
    if x is None:
        return None
    if _is_tensor(x):
        assert x.is_cuda, "tensors passed to colosseum_b200 must live on the GPU"
        tdt = {np.float32: torch.float32, np.float64: torch.float64, np.int32: torch.int32, np.uint8: torch.uint8}[np_dtype]
        return x.to(tdt).contiguous()
    a = np.ascontiguousarray(x, dtype=np_dtype)
    return torch.from_numpy(a).cuda()


def _result(t, as_numpy):
    return t.cpu().numpy() if as_numpy else t


def _vdtype(precision):
    torch = _torch()
    return torch.float64 if precision == "f64" else torch.float32


def _scratch(nbytes):
    torch = _torch()
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device="cuda")


# ------------------------------------------------------------------------------------------------ discounted
def _solve_discounted(T, R, pi, gamma, epsilon, max_abs_value, precision, max_iter=DP_MAX_ITERATION, sweep_order=None):
    precision = precision or _PRECISION
    sweep_order = sweep_order or _SWEEP_ORDER
    as_numpy = not _is_tensor(T)
    torch = _torch()
    Td, Rd, pid = to_device(T), to_device(R), to_device(pi)
    batched = Td.dim() == 4
    if not batched:
        Td, Rd = Td[None], Rd[None]
        pid = None if pid is None else pid[None]
    B, S, A, S2 = Td.shape
    assert S == S2 and tuple(Rd.shape) == (B, S, A)
    assert pid is None or tuple(pid.shape) == (B, S, A)
    vd = _vdtype(precision)
    Q = torch.empty((B, S, A), dtype=vd, device="cuda")
    V = torch.empty((B, S), dtype=vd, device="cuda")
    lib = _cabi.lib()
    f64 = precision == "f64"
    work = _scratch(lib.colo_solve_work_bytes(B, S, int(f64)))
    iters = (C.c_longlong * B)()
    fold = FOLD_PI if pid is not None else FOLD_MAX
    # gamma is cast to float32 first, as the reference does (infinite_horizon.py:127,171)
    g = float(np.float32(gamma))
    if sweep_order == "gauss_seidel":
        it_d = torch.zeros(B, dtype=torch.int64, device="cuda")
        st_d = torch.zeros(B, dtype=torch.int32, device="cuda")
        fn = lib.colo_solve_discounted_gs_f64acc if f64 else lib.colo_solve_discounted_gs_f32
        rc = fn(_cabi.ptr(Td), _cabi.ptr(Rd), _cabi.ptr(pid), B, S, A, g, float(epsilon),
                float(max_abs_value) if max_abs_value is not None else 0.0, int(max_iter), fold,
                _cabi.ptr(Q), _cabi.ptr(V), _cabi.ptr(it_d), _cabi.ptr(st_d), _cabi.current_stream())
        _cabi.check(rc, "colo_solve_discounted_gs")
        st_h = st_d.cpu().numpy()
        for b, v in enumerate(it_d.cpu().numpy()):
            iters[b] = int(v)
        rc = _cabi.OVERFLOW if (st_h == _cabi.OVERFLOW).any() else (_cabi.MAX_ITER if (st_h == _cabi.MAX_ITER).any() else 0)
    else:
        fn = lib.colo_solve_discounted_f64acc if f64 else lib.colo_solve_discounted_f32
        rc = fn(_cabi.ptr(Td), _cabi.ptr(Rd), _cabi.ptr(pid), B, S, A, g, float(epsilon),
                float(max_abs_value) if max_abs_value is not None else 0.0, int(max_iter), fold,
                _cabi.ptr(Q), _cabi.ptr(V), iters, _cabi.ptr(work), _cabi.current_stream())
        _cabi.check(rc, "colo_solve_discounted")
    if rc == _cabi.OVERFLOW:
        return None
    if rc == _cabi.MAX_ITER:
        raise DynamicProgrammingMaxIterationExceeded()
    if not batched:
        Q, V = Q[0], V[0]
    out = _result(Q, as_numpy), _result(V, as_numpy)
    _solve_discounted.last_iterations = list(iters)
    return out


_solve_discounted.last_iterations = []


def last_iterations():
    """sweeps run by each instance of the most recent discounted solve (diagnostics / bench)."""
    return list(_solve_discounted.last_iterations)


def discounted_value_iteration(T, R, gamma=0.99, epsilon=1e-3, max_abs_value=None,
                               sparse_n_states_threshold=300 * 3 * 300, sparse_nnz_per_threshold=0.2, *,
                               precision=None, sweep_order=None):
    """colosseum/dynamic_programming/infinite_horizon.py:14-44.  Returns (Q[S,A], V[S]) or None on overflow.

    The two `sparse_*` arguments select a pydata-sparse code path in the reference; the GPU path streams the dense
    tensor at HBM speed for every size, so they are accepted and ignored."""
    return _solve_discounted(T, R, None, gamma, epsilon, max_abs_value, precision, sweep_order=sweep_order)


def discounted_policy_evaluation(T, R, pi, gamma=0.99, epsilon=1e-7, sparse_n_states_threshold=200,
                                 sparse_nnz_per_threshold=0.2, *, precision=None, sweep_order=None):
    """colosseum/dynamic_programming/infinite_horizon.py:47-64.  Returns (Q[S,A], V[S])."""
    return _solve_discounted(T, R, pi, gamma, epsilon, None, precision, sweep_order=sweep_order)


def discounted_policy_iteration(T, R, gamma=0.99, epsilon=1e-7, *, precision=None):
    """colosseum/dynamic_programming/infinite_horizon.py:208-219.  Returns (Q, V, pi)."""
    S, A = R.shape[-2], R.shape[-1]
    rng = np.random.RandomState(ARGMAX_SEED)
    Q = rng.rand(S, A)
    pi = get_policy_from_q_values(Q, True)
    for _ in range(DP_MAX_ITERATION):
        old_pi = pi.copy()
        Q, V = discounted_policy_evaluation(T, R, pi, gamma, epsilon, precision=precision)
        Qh = Q.cpu().numpy() if _is_tensor(Q) else Q
        pi = get_policy_from_q_values(Qh, True)
        if (pi != old_pi).sum() == 0:
            return Q, V, pi
    raise DynamicProgrammingMaxIterationExceeded()


def extended_value_iteration(T, estimated_rewards, beta_r, beta_p, r_max, epsilon=1e-3, *, precision=None,
                             max_iter=DP_MAX_ITERATION, return_iterations=False):
    """colosseum/dynamic_programming/infinite_horizon.py:67-118 (UCRL2's optimistic value iteration, with `_max_proba`
    :222-251).  Returns (span of the value function, Q[S,A], V[S]), or None when the iteration cap is reached.
    `beta_p` may be [S,A], [S,A,1] or [S,A,S]: like the reference (:230) only element 0 of the last axis is used."""
    precision = precision or _PRECISION
    as_numpy = not _is_tensor(T)
    torch = _torch()
    Td, Rd = to_device(T), to_device(estimated_rewards)
    S, A, _ = Td.shape
    br = to_device(beta_r, np.float64).reshape(S, A)
    bp = beta_p
    if getattr(bp, "ndim", 2) == 3 or (hasattr(bp, "dim") and bp.dim() == 3):
        bp = bp[..., 0]
    bp = to_device(np.ascontiguousarray(bp) if not _is_tensor(bp) else bp, np.float64).reshape(S, A)
    f64 = precision == "f64"
    vd = _vdtype(precision)
    Q = torch.empty((S, A), dtype=vd, device="cuda")
    V = torch.empty(S, dtype=vd, device="cuda")
    lib = _cabi.lib()
    work = _scratch(lib.colo_extended_vi_work_bytes(S, int(f64)))
    out = (C.c_double * 2)()
    fn = lib.colo_extended_vi_f64acc if f64 else lib.colo_extended_vi_f32
    rc = fn(_cabi.ptr(Td), _cabi.ptr(Rd), _cabi.ptr(br), _cabi.ptr(bp), S, A, float(r_max), float(epsilon), int(max_iter),
            _cabi.ptr(Q), _cabi.ptr(V), out, _cabi.ptr(work), _cabi.current_stream())
    _cabi.check(rc, "colo_extended_vi")
    if rc == _cabi.MAX_ITER:
        return None
    res = (float(out[0]), _result(Q, as_numpy), _result(V, as_numpy))
    return res + (int(out[1]),) if return_iterations else res


# ------------------------------------------------------------------------------------------------ episodic
def _episodic(H, T, R, policy, max_value, precision):
    precision = precision or _PRECISION
    as_numpy = not _is_tensor(T)
    torch = _torch()
    H = int(H)
    Td, Rd, pid = to_device(T), to_device(R), to_device(policy)
    batched = Td.dim() == 4
    if not batched:
        Td, Rd = Td[None], Rd[None]
        pid = None if pid is None else pid[None]
    B, S, A, _ = Td.shape
    if pid is not None and pid.shape[1] == H + 1:
        # a policy derived from the (H+1)-row Q of episodic_value_iteration (PSRLEpisodic.current_optimal_stochastic_
        # policy, posterior_sampling.py:76-80): the reference only ever reads rows 0..H-1 (finite_horizon.py:36-40)
        pid = pid[:, :H].contiguous()
    assert pid is None or tuple(pid.shape) == (B, H, S, A), "policy must be [H,S,A]"
    vd = _vdtype(precision)
    Q = torch.empty((B, H + 1, S, A), dtype=vd, device="cuda")
    V = torch.empty((B, H + 1, S), dtype=vd, device="cuda")
    lib = _cabi.lib()
    fn = lib.colo_episodic_f64acc if precision == "f64" else lib.colo_episodic_f32
    rc = fn(_cabi.ptr(Td), _cabi.ptr(Rd), _cabi.ptr(pid), B, S, A, H, FOLD_PI if pid is not None else FOLD_MAX,
            float(max_value) if max_value is not None else 0.0, _cabi.ptr(Q), _cabi.ptr(V), _cabi.current_stream())
    _cabi.check(rc, "colo_episodic")
    if rc == _cabi.OVERFLOW:
        return None
    if not batched:
        Q, V = Q[0], V[0]
    return _result(Q, as_numpy), _result(V, as_numpy)


def episodic_value_iteration(H, T, R, max_value=None, *, precision=None):
    """colosseum/dynamic_programming/finite_horizon.py:11-26.  Returns (Q[H+1,S,A], V[H+1,S]) or None."""
    return _episodic(H, T, R, None, max_value, precision)


def episodic_policy_evaluation(H, T, R, policy, *, precision=None):
    """colosseum/dynamic_programming/finite_horizon.py:29-42.  policy is [H,S,A]."""
    return _episodic(H, T, R, policy, None, precision)


# ------------------------------------------------------------------------------------------------ single backups
def bellman_backup(T, R, V, gamma=0.99, pi=None, fold=None, *, precision=None, return_residual=False):
    """One synchronous sweep Q = R + gamma T V, V' = fold(Q) (the body of infinite_horizon.py:131-135).
    Building block for callers that own their iteration (agents, sharded solvers, benchmarks)."""
    precision = precision or _PRECISION
    as_numpy = not _is_tensor(T)
    torch = _torch()
    f64 = precision == "f64"
    Td, Rd, pid = to_device(T), to_device(R), to_device(pi)
    Vd = to_device(V, np.float64 if f64 else np.float32)
    batched = Td.dim() == 4
    if not batched:
        Td = Td[None]
        Rd = None if Rd is None else Rd[None]
        Vd = Vd[None]
        pid = None if pid is None else pid[None]
    B, S, A, _ = Td.shape
    vd = _vdtype(precision)
    Q = torch.empty((B, S, A), dtype=vd, device="cuda")
    Vn = torch.empty((B, S), dtype=vd, device="cuda")
    resid = torch.zeros(B, dtype=torch.int64 if f64 else torch.int32, device="cuda")
    a = _cabi.BackupArgs()
    a.T, a.R, a.pi = _cabi.ptr(Td), _cabi.ptr(Rd), _cabi.ptr(pid)
    a.V_in, a.V_out, a.Q = _cabi.ptr(Vd), _cabi.ptr(Vn), _cabi.ptr(Q)
    a.t_stride, a.r_stride, a.pi_stride = S * A * S, S * A, S * A
    a.v_in_stride, a.v_out_stride, a.q_stride = S, S, S * A
    a.B, a.S, a.A = B, S, A
    a.fold = fold if fold is not None else (FOLD_PI if pid is not None else FOLD_MAX)
    a.gamma = float(gamma)
    a.resid = _cabi.ptr(resid)
    a.row0, a.nrows = 0, S
    lib = _cabi.lib()
    rc = (lib.colo_backup_f64acc if f64 else lib.colo_backup_f32)(C.byref(a), _cabi.current_stream())
    _cabi.check(rc, "colo_backup")
    if not batched:
        Q, Vn = Q[0], Vn[0]
    out = (_result(Q, as_numpy), _result(Vn, as_numpy))
    if return_residual:
        r = resid.view(torch.float64 if f64 else torch.float32)
        out = out + (_result(r, as_numpy),)
    return out


# ------------------------------------------------------------------------------------------------ argmax helpers
def _argmax_rows(Q2, rng):
    """random tie-broken argmax per row (colosseum/dynamic_programming/utils.py:12-72)."""
    best = Q2.max(-1, keepdims=True)
    out = np.empty(Q2.shape[0], np.int32)
    ties = Q2 == best
    n_ties = ties.sum(-1)
    first = ties.argmax(-1)
    out[:] = first
    for r in np.nonzero(n_ties > 1)[0]:
        out[r] = rng.choice(np.nonzero(ties[r])[0])
    return out


def get_policy_from_q_values(Q, stochastic_form=False):
    """colosseum/dynamic_programming/utils.py:75-100: deterministic policy from Q; ties are broken at random
    under a fixed seed (ARGMAX_SEED = 42).  NOTE: the reference draws its tie-breaks from numba's private
    Mersenne-Twister stream; this host helper uses numpy's RandomState(42), so the CHOICE among exactly tied
    actions can differ -- every choice is an argmax, and the value functions are unaffected."""
    Qh = Q.cpu().numpy() if _is_tensor(Q) else np.asarray(Q)
    rng = np.random.RandomState(ARGMAX_SEED)
    lead = Qh.shape[:-1]
    idx = _argmax_rows(Qh.reshape(-1, Qh.shape[-1]), rng).reshape(lead)
    if not stochastic_form:
        return idx.astype(np.int32)
    X = np.zeros(Qh.shape, np.float32)
    np.put_along_axis(X, idx[..., None].astype(np.int64), 1.0, axis=-1)
    return X


# ------------------------------------------------------------------------------------------------ resident sweeper
class BatchedValueIteration:
    """Device-resident synchronous value iteration over a batch of MDPs: T, R stay in HBM, V ping-pongs, and
    `sweep()` is exactly one launch of the backup kernel (no allocation, no host sync).  This is what bench.py
    times for config C4/C5 and what a model-based agent that re-plans every episode would hold on to."""

    def __init__(self, T, R, gamma=0.99, pi=None, precision=None, row0=0, S_total=None):
        torch = _torch()
        self.precision = precision or _PRECISION
        self.f64 = self.precision == "f64"
        self.T, self.R, self.pi = to_device(T), to_device(R), to_device(pi)
        if self.T.dim() == 3:
            self.T, self.R = self.T[None], self.R[None]
            self.pi = None if self.pi is None else self.pi[None]
        B, nrows, A, S = self.T.shape
        self.B, self.nrows, self.A, self.S = B, nrows, A, S
        self.row0 = int(row0)
        assert S == (S_total or S) and self.row0 + nrows <= S
        vd = _vdtype(self.precision)
        self.V = [torch.zeros((B, S), dtype=vd, device="cuda") for _ in range(2)]
        self.Q = torch.zeros((B, nrows, A), dtype=vd, device="cuda")
        self.resid = torch.zeros(B, dtype=torch.int64 if self.f64 else torch.int32, device="cuda")
        self.cur = 0
        a = _cabi.BackupArgs()
        a.T, a.R, a.pi, a.Q = _cabi.ptr(self.T), _cabi.ptr(self.R), _cabi.ptr(self.pi), _cabi.ptr(self.Q)
        a.t_stride, a.r_stride, a.pi_stride = nrows * A * S, nrows * A, nrows * A
        a.v_in_stride, a.v_out_stride, a.q_stride = S, S, nrows * A
        a.B, a.S, a.A = B, S, A
        a.fold = FOLD_PI if self.pi is not None else FOLD_MAX
        a.gamma = float(np.float32(gamma))
        a.resid = _cabi.ptr(self.resid)
        a.row0, a.nrows = self.row0, nrows
        self.args = a
        lib = _cabi.lib()
        self._fn = lib.colo_backup_f64acc if self.f64 else lib.colo_backup_f32
        self.sweeps = 0

    @property
    def values(self):
        return self.V[self.cur]

    def set_peers(self, peer_ptr_tensors):
        """row-sharded mode: `peer_ptr_tensors[i]` is a CUDA int64 tensor holding every rank's pointer to its V[i]
        buffer (i = 0, 1: the two ping-pong buffers); the sweep then also stores its rows into the peers."""
        self._peers = peer_ptr_tensors

    def sweep(self, n=1, store_q=True):
        a = self.args
        a.Q = _cabi.ptr(self.Q) if store_q else None
        stream = _cabi.current_stream()
        peers = getattr(self, "_peers", None)
        for _ in range(n):
            nxt = 1 - self.cur
            a.V_in, a.V_out = _cabi.ptr(self.V[self.cur]), _cabi.ptr(self.V[nxt])
            if peers is not None:
                a.V_out_peers, a.n_peers = _cabi.ptr(peers[nxt]), int(peers[nxt].numel())
            _cabi.check(self._fn(C.byref(a), stream), "colo_backup")
            self.cur = nxt
            self.sweeps += 1

    def residual(self):
        """max|dV| per instance accumulated since the last call (device tensor); resets the accumulator."""
        torch = _torch()
        r = self.resid.view(torch.float64 if self.f64 else torch.float32).clone()
        self.resid.zero_()
        return r
