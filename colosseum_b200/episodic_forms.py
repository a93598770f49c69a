"""B200 twins of the reference's episodic tensor builders (colosseum/mdp/utils/mdp_creation.py:98-176) and of
`EpisodicMDP.reachable_states` (colosseum/mdp/base_finite.py:138-150).

The reference builds these with Python loops over a networkx graph; here they are scatter kernels over the dense T
already resident in HBM (csrc/builders.cu).  Instead of a NextStateSampler + node_to_index the functions take the
start distribution as index / probability arrays (what `MDPTables` holds).  Results are CUDA tensors when T is a
CUDA tensor, numpy arrays otherwise -- same convention as colosseum_b200.dynamic_programming.
"""
import numpy as np

from . import _cabi
from .dynamic_programming import _is_tensor, _result, _torch, to_device


def _start(start_idx, start_prob):
    torch = _torch()
    si = torch.from_numpy(np.ascontiguousarray(start_idx, np.int32)).cuda()
    sp = torch.from_numpy(np.ascontiguousarray(start_prob, np.float64)).cuda()
    assert si.numel() == sp.numel() and si.numel() > 0
    return si, sp


def get_episodic_transition_matrix_and_rewards(H, T, R, start_idx, start_prob, return_reach=False):
    """mdp_creation.py:98-128.  Returns (T_epi f32[H,S,A,S], R_epi f32[H,S,A]) [+ reach bool[H,S]]."""
    torch = _torch()
    as_numpy = not _is_tensor(T)
    Td, Rd = to_device(T), to_device(R)
    S, A, _ = Td.shape
    H = int(H)
    si, sp = _start(start_idx, start_prob)
    T_epi = torch.empty((H, S, A, S), dtype=torch.float32, device="cuda")
    R_epi = torch.empty((H, S, A), dtype=torch.float32, device="cuda")
    reach = torch.empty((H, S), dtype=torch.uint8, device="cuda")
    rc = _cabi.lib().colo_build_episodic_tensor(_cabi.ptr(Td), _cabi.ptr(Rd), _cabi.ptr(si), _cabi.ptr(sp), si.numel(), H, S, A,
                                           _cabi.ptr(T_epi), _cabi.ptr(R_epi), _cabi.ptr(reach), _cabi.current_stream())
    _cabi.check(rc, "colo_build_episodic_tensor")
    out = (_result(T_epi, as_numpy), _result(R_epi, as_numpy))
    if return_reach:
        out = out + (_result(reach.bool(), as_numpy),)
    return out


def reachable_states(H, T, start_idx, start_prob):
    """base_finite.py:138-150 as a sorted list of (h, state index) pairs."""
    torch = _torch()
    Td = to_device(T)
    S, A, _ = Td.shape
    H = int(H)
    si, sp = _start(start_idx, start_prob)
    # the reachability pass of the builder without keeping T_epi would need its own kernel; T_epi is H*S*A*S*4 B,
    # affordable for every benchmark instance (<= 45 MB) -- reuse the builder
    T_epi = torch.empty((H, S, A, S), dtype=torch.float32, device="cuda")
    reach = torch.empty((H, S), dtype=torch.uint8, device="cuda")
    rc = _cabi.lib().colo_build_episodic_tensor(_cabi.ptr(Td), None, _cabi.ptr(si), _cabi.ptr(sp), si.numel(), H, S, A,
                                           _cabi.ptr(T_epi), None, _cabi.ptr(reach), _cabi.current_stream())
    _cabi.check(rc, "colo_build_episodic_tensor")
    hs = torch.nonzero(reach).cpu().numpy()
    return [(int(h), int(s)) for h, s in hs]


def get_continuous_form_episodic_transition_matrix_and_rewards(H, T, R, start_idx, start_prob, nodes=None):
    """mdp_creation.py:131-176.  `nodes`: the (h, s) pairs in the order the rows/columns of T_cf should have (pass
    the reference's `mdp.reachable_states` to reproduce `mdp.T_cf` bit for bit); default = all reachable pairs sorted
    by (h, s).  Returns (T_cf f32[n,A,n], R_cf f32[n,A])."""
    torch = _torch()
    as_numpy = not _is_tensor(T)
    Td, Rd = to_device(T), to_device(R)
    S, A, _ = Td.shape
    H = int(H)
    if nodes is None:
        nodes = reachable_states(H, Td, start_idx, start_prob)
    hs = np.asarray(list(nodes), np.int64).reshape(-1, 2)
    n = len(hs)
    pos = -np.ones((H, S), np.int32)
    pos[hs[:, 0], hs[:, 1]] = np.arange(n, dtype=np.int32)
    assert (pos >= 0).sum() == n, "duplicate (h, s) pairs in nodes"
    si, sp = _start(start_idx, start_prob)
    nh = torch.from_numpy(hs[:, 0].astype(np.int32)).cuda()
    ns = torch.from_numpy(hs[:, 1].astype(np.int32)).cuda()
    posd = torch.from_numpy(pos).cuda()
    T_cf = torch.empty((n, A, n), dtype=torch.float32, device="cuda")
    R_cf = torch.empty((n, A), dtype=torch.float32, device="cuda")
    flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    rc = _cabi.lib().colo_build_continuous_form(_cabi.ptr(Td), _cabi.ptr(Rd), _cabi.ptr(nh), _cabi.ptr(ns), n,
                                                _cabi.ptr(posd), _cabi.ptr(si), _cabi.ptr(sp), si.numel(), H, S, A,
                                                _cabi.ptr(T_cf), _cabi.ptr(R_cf), _cabi.ptr(flag), _cabi.current_stream())
    _cabi.check(rc, "colo_build_continuous_form")
    # the reference asserts np.isclose(T_cf.sum(-1), 1).all() (mdp_creation.py:174)
    assert int(flag.item()) == 0, "a positive-probability successor is missing from the node list"
    return _result(T_cf, as_numpy), _result(R_cf, as_numpy)


def continuous_form_optimal_values(H, T, R, start_idx, start_prob, nodes=None, gamma=0.99, epsilon=None,
                                   precision="f64", max_evaluations=200):
    """`mdp.optimal_value_continuous_form[1]` (mdp/base_finite.py:167-178: discounted_value_iteration on T_cf, R_cf) at
    the FIXED POINT, computed from the structure of the continuous form instead of sweeping its n x A x n tensor
    (colo_continuous_form_values_*: the fixed point depends on one scalar -- the value the last layer jumps to --
    found by a handful of backward inductions over the original T).  `nodes` as in
    get_continuous_form_episodic_transition_matrix_and_rewards; returns V_cf [n] in that node order."""
    import ctypes as C

    torch = _torch()
    as_numpy = not _is_tensor(T)
    Td, Rd = to_device(T), to_device(R)
    S, A, _ = Td.shape
    H = int(H)
    if nodes is None:
        nodes = reachable_states(H, Td, start_idx, start_prob)
    hs = np.asarray(list(nodes), np.int64).reshape(-1, 2)
    si = np.ascontiguousarray(start_idx, np.int64)
    assert si.max() < len(hs), "a start index beyond the node list (mdp_creation.py:168 writes column node_to_index[start])"
    f64 = precision == "f64"
    eps = float(epsilon if epsilon is not None else (1e-9 if f64 else 1e-5))
    # the node sitting at list position start_k (sic: the reference uses the ORIGINAL state index as a column index)
    ph = torch.from_numpy(hs[si, 0].astype(np.int32)).cuda()
    ps = torch.from_numpy(hs[si, 1].astype(np.int32)).cuda()
    p32 = torch.from_numpy(np.ascontiguousarray(start_prob, np.float64).astype(np.float32)).cuda()
    V = torch.empty((H, S), dtype=torch.float64 if f64 else torch.float32, device="cuda")
    out = (C.c_double * 2)()
    lib = _cabi.lib()
    fn = lib.colo_continuous_form_values_f64acc if f64 else lib.colo_continuous_form_values_f32
    rc = fn(_cabi.ptr(Td), _cabi.ptr(Rd), S, A, H, float(np.float32(gamma)), _cabi.ptr(ph), _cabi.ptr(ps), _cabi.ptr(p32),
            len(si), eps, int(max_evaluations), _cabi.ptr(V), out, _cabi.current_stream())
    _cabi.check(rc, "colo_continuous_form_values")
    if rc == _cabi.MAX_ITER:
        from .dynamic_programming import DynamicProgrammingMaxIterationExceeded

        raise DynamicProgrammingMaxIterationExceeded()
    continuous_form_optimal_values.last_evaluations = int(out[1])
    V_cf = V[torch.from_numpy(hs[:, 0]).cuda(), torch.from_numpy(hs[:, 1]).cuda()]
    return _result(V_cf, as_numpy)


continuous_form_optimal_values.last_evaluations = 0
