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
