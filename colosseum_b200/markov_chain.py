"""B200 twin of the average-reward helpers of `colosseum/mdp/utils/markov_chain.py` (get_average_reward :12-31,
get_average_rewards :34-41, get_transition_probabilities :44-51, get_stationary_distribution :64-137), the functions
behind `BaseMDP.{optimal,worst,random}_average_reward` (mdp/base.py:895-941) and the continuous-MDP regret of
`MDPLoop` (experiment/agent_mdp_interaction.py:518-578).

The reference finds the recurrent classes with networkx and solves each with GTH elimination / ARPACK.  Here the
chain of the policy is built in one pass over T and its stationary distribution is the limit of the START distribution
under the lazy chain (I + P)/2, computed on the GPU by repeated squaring in fp64.  Which start vector: the
reference's own rule, restated on the host with the same networkx calls -- one recurrent class: that class's
stationary distribution whatever the start; several: every start state's whole mass goes to the FIRST class (in
networkx's attracting-component order) it can reach (:113-126)."""
import ctypes as C

import numpy as np

from . import _cabi
from .dynamic_programming import DynamicProgrammingMaxIterationExceeded, _is_tensor, _result, _scratch, _torch, to_device


def _chain(T, R, policy):
    torch = _torch()
    Td, pid = to_device(T), to_device(policy)
    Rd = None if R is None else to_device(R)
    S, A, _ = Td.shape
    P = torch.empty((S, S), dtype=torch.float32, device="cuda")
    r = None if Rd is None else torch.empty(S, dtype=torch.float32, device="cuda")
    rc = _cabi.lib().colo_policy_chain(_cabi.ptr(Td), _cabi.ptr(Rd), _cabi.ptr(pid), S, A, _cabi.ptr(P), _cabi.ptr(r),
                                       _cabi.current_stream())
    _cabi.check(rc, "colo_policy_chain")
    return P, r


def get_average_rewards(R, policy):
    """markov_chain.py:34-41: expected reward per state under the policy."""
    as_numpy = not _is_tensor(R)
    Rd, pid = to_device(R), to_device(policy)
    return _result((Rd * pid).sum(-1), as_numpy)


def get_transition_probabilities(T, policy):
    """markov_chain.py:44-51: the transition matrix of the Markov chain the policy yields, f32 [S,S]."""
    P, _ = _chain(T, None, policy)
    return _result(P, not _is_tensor(T))


def recurrent_class_weights(tps_host, starting_states_and_probs, sparse_threshold_size=500 * 500):
    """The reference's treatment of multichain policies, restated on the host (markov_chain.py:90-133): the recurrent
    classes are networkx's attracting components of the chain's graph, in networkx's order; with ONE class the start
    distribution is irrelevant; with several, each start state's WHOLE mass goes to the first class (in that order) it
    can reach.  Returns the start vector x0 f64[S] that encodes those weights as mass on one representative state of
    each class -- started inside a closed class the chain never leaves it, so lim x0 L^n = sum_c w_c pi_c."""
    import networkx as nx

    S = len(tps_host)
    if tps_host.size > sparse_threshold_size:
        from scipy.sparse import coo_matrix

        G = nx.DiGraph(coo_matrix(tps_host))
    else:
        G = nx.DiGraph(tps_host)
    classes = list(map(tuple, nx.attracting_components(G)))
    x0 = np.zeros(S, np.float64)
    if len(classes) == 1:
        x0[classes[0][0]] = 1.0
        return x0, 1
    for ss, p in starting_states_and_probs:  # None raises TypeError here, as in the reference
        for c in classes:
            if nx.has_path(G, int(ss), c[0]):
                x0[c[0]] += float(p)
                break
    return x0, len(classes)


def get_stationary_distribution(tps, starting_states_and_probs=None, sparse_threshold_size=500 * 500, *, tol=1e-13,
                                max_squarings=200):
    """markov_chain.py:64-137.  `starting_states_and_probs`: iterable of (state index, probability), needed (as in the
    reference) only when the chain has several recurrent classes.  The class structure and the reference's
    first-reachable-class rule are resolved on the host (`recurrent_class_weights`); the stationary distributions
    themselves are the limit of the resulting start vector under the lazy chain (I + P)/2, computed on the GPU by
    repeated squaring in fp64 (robust for nearly reducible chains)."""
    torch = _torch()
    as_numpy = not _is_tensor(tps)
    P = to_device(tps)
    S = int(P.shape[0])
    host = np.asarray(tps) if as_numpy else P.cpu().numpy()
    x0, n_classes = recurrent_class_weights(host, starting_states_and_probs, sparse_threshold_size or 500 * 500)
    get_stationary_distribution.last_classes = n_classes
    x0d = torch.from_numpy(x0).cuda()
    lib = _cabi.lib()
    x = torch.empty(S, dtype=torch.float64, device="cuda")
    work = _scratch(lib.colo_stationary_distribution_work_bytes(S))
    k = (C.c_int * 1)()
    rc = lib.colo_stationary_distribution_f64(_cabi.ptr(P), S, _cabi.ptr(x0d), float(tol), int(max_squarings), _cabi.ptr(x),
                                              k, _cabi.ptr(work), _cabi.current_stream())
    _cabi.check(rc, "colo_stationary_distribution_f64")
    if rc == _cabi.MAX_ITER:
        raise DynamicProgrammingMaxIterationExceeded()
    get_stationary_distribution.last_iterations = int(k[0])
    return _result(x, as_numpy)


def power_iteration(tps, x0, tol=1e-10, max_iter=int(1e6)):
    """x <- x (I + P)/2, rescaled to unit sum, from x0 until max|dx| < tol -- on the backup kernels (compressed rows on
    chip when the chain is sparse).  For well-mixing chains only (see colo_power_iteration_f64)."""
    torch = _torch()
    as_numpy = not _is_tensor(tps)
    P = to_device(tps)
    S = int(P.shape[0])
    x0d = to_device(x0, np.float64)
    lib = _cabi.lib()
    M = torch.empty((S, S), dtype=torch.float32, device="cuda")
    _cabi.check(lib.colo_lazy_transpose(_cabi.ptr(P), S, _cabi.ptr(M), _cabi.current_stream()), "colo_lazy_transpose")
    x = torch.empty(S, dtype=torch.float64, device="cuda")
    work = _scratch(lib.colo_power_iteration_work_bytes(S))
    iters = (C.c_longlong * 1)()
    rc = lib.colo_power_iteration_f64(_cabi.ptr(M), S, _cabi.ptr(x0d), float(tol), int(max_iter), _cabi.ptr(x), iters,
                                      _cabi.ptr(work), _cabi.current_stream())
    _cabi.check(rc, "colo_power_iteration_f64")
    if rc == _cabi.MAX_ITER:
        raise DynamicProgrammingMaxIterationExceeded()
    return _result(x, as_numpy), int(iters[0])


get_stationary_distribution.last_iterations = 0
get_stationary_distribution.last_classes = 0


def get_average_reward(T, R, policy, next_states_and_probs=None, sparse_threshold_size=None, *, tol=1e-13):
    """markov_chain.py:12-31: the expected time-average reward of `policy` (stochastic form [S,A])."""
    pol = policy.cpu().numpy() if _is_tensor(policy) else np.asarray(policy)
    assert np.isclose(pol.sum(-1), 1).all(), "the policy specification is incorrect."
    P, r = _chain(T, R, policy)
    sd = get_stationary_distribution(P, next_states_and_probs, tol=tol)
    return float((r.double() * sd).sum().item())


def get_average_reward_batched(T, R, policies, start_states=None, *, tol=1e-13, max_squarings=200,
                               max_work_bytes=8 << 30):
    """markov_chain.py:12-31 for a batch of policies f32 [B,S,A] on one MDP: f64 [B] average rewards, one batched
    repeated-squaring solve per chunk (colo_average_rewards_f64).  Policies whose chain has several recurrent classes
    (the rows of the limiting matrix disagree) go through `get_average_reward` one by one with the reference's
    first-reachable-class rule, which needs `start_states` (i [B]: where each policy's loop currently is)."""
    torch = _torch()
    Td, Rd = to_device(T), to_device(R)
    pid = to_device(policies)
    B, S, A = pid.shape
    lib = _cabi.lib()
    chunk = max(1, min(65535, int(max_work_bytes // max(1, lib.colo_average_rewards_work_bytes(1, S)))))
    ar = torch.empty(B, dtype=torch.float64, device="cuda")
    multi = torch.empty(B, dtype=torch.int32, device="cuda")
    for lo in range(0, B, chunk):
        n = min(chunk, B - lo)
        work = _scratch(lib.colo_average_rewards_work_bytes(n, S))
        rc = lib.colo_average_rewards_f64(_cabi.ptr(Td), _cabi.ptr(Rd), _cabi.ptr(pid[lo:lo + n]), n, S, A, float(tol),
                                          int(max_squarings), _cabi.ptr(ar[lo:lo + n]), _cabi.ptr(multi[lo:lo + n]), None,
                                          _cabi.ptr(work), _cabi.current_stream())
        _cabi.check(rc, "colo_average_rewards_f64")
        if rc == _cabi.MAX_ITER:
            raise DynamicProgrammingMaxIterationExceeded()
    out = ar.cpu().numpy()
    flagged = np.nonzero(multi.cpu().numpy())[0]
    get_average_reward_batched.last_multichain = len(flagged)
    for b in flagged:
        start = None if start_states is None else [(int(start_states[b]), 1.0)]
        out[b] = get_average_reward(Td, Rd, pid[b], start, tol=tol)
    return out


get_average_reward_batched.last_multichain = 0
