"""Python face of the CPU oracle (oracle/colo_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import this
module.  The product package (colosseum_b200/) never does: it has no CPU path at all.

Parity status: PINNED against the unmodified reference -- see tests/test_oracle_golden.py and
tests/golden/make_golden.py.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libcolo_oracle.so")

OK, OVERFLOW, MAX_ITER, NEEDS_RESET = 0, 1, 2, 3
FOLD_MAX, FOLD_PI, FOLD_MIN = 0, 1, 2


def build(force=False):
    if force or not os.path.isfile(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "colo_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_value_norm_f64.restype = C.c_double
        _lib.orc_gaps_f64.restype = C.c_double
        _lib.orc_diameter_target_ref_f32.restype = C.c_float
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, np.float32)


class Tables(C.Structure):
    """mirror of `orc_tables` / the product's `colo_mdp_tables` (host pointers here)."""

    _fields_ = [
        ("S", C.c_int), ("A", C.c_int), ("H", C.c_int), ("ld", C.c_int),
        ("cdf", C.c_void_p),
        ("succ_cum", C.c_void_p), ("succ_idx", C.c_void_p), ("succ_len", C.c_void_p), ("Ksucc", C.c_int),
        ("rew_cls_sas", C.c_void_p), ("rew_cls_sa", C.c_void_p), ("rew_cls_succ", C.c_void_p),
        ("rew_q", C.c_void_p), ("n_cls", C.c_int), ("nq", C.c_int),
        ("rmin", C.c_float), ("rmax", C.c_float),
        ("start_cum", C.c_void_p), ("start_idx", C.c_void_p), ("n_start", C.c_int),
    ]


# ---------------------------------------------------------------------------------------------- DP
def discounted_gs_f32(T, R, pi=None, gamma=0.99, eps=1e-3, max_abs=None, max_iter=int(1e6)):
    """the reference iterate (infinite_horizon.py:121-142 / :167-184): fp32, in place.  Returns (Q,V,iters)|None."""
    T, R, pi = _f32(T), _f32(R), _f32(pi)
    S, A, _ = T.shape
    Q = np.zeros((S, A), np.float32)
    V = np.zeros(S, np.float32)
    it = C.c_longlong(0)
    rc = lib().orc_discounted_gs_f32(_p(T), _p(R), _p(pi), S, A, C.c_float(gamma), C.c_float(eps),
                                     C.c_float(max_abs or 0.0), C.c_longlong(max_iter), _p(Q), _p(V), C.byref(it))
    if rc == OVERFLOW:
        return None
    if rc == MAX_ITER:
        raise RuntimeError("max iterations exceeded")
    return Q, V, it.value


def discounted_f64(T, R, pi=None, gamma=0.99, tol=1e-12, max_abs=None, max_iter=int(1e6), fold=FOLD_MAX,
                   gauss_seidel=False):
    """fixed-point oracle in fp64.  Returns (Q,V,iters) | None."""
    T, R, pi = _f32(T), _f32(R), _f32(pi)
    S, A, _ = T.shape
    if pi is not None:
        fold = FOLD_PI
    Q = np.zeros((S, A), np.float64)
    V = np.zeros(S, np.float64)
    it = C.c_longlong(0)
    rc = lib().orc_discounted_f64(_p(T), _p(R), _p(pi), S, A, C.c_double(gamma), C.c_double(tol),
                                  C.c_double(max_abs or 0.0), C.c_longlong(max_iter), fold, int(gauss_seidel),
                                  _p(Q), _p(V), C.byref(it))
    if rc == OVERFLOW:
        return None
    if rc == MAX_ITER:
        raise RuntimeError("max iterations exceeded")
    return Q, V, it.value


def jacobi_sweeps_f64(T, R, V0, n, gamma=0.99, pi=None, fold=FOLD_MAX):
    """n synchronous sweeps from V0 in fp64 (numpy) -- checks single backups / fixed sweep counts."""
    T64 = np.asarray(T, np.float64)
    V = np.asarray(V0, np.float64).copy()
    Q = None
    for _ in range(n):
        Q = (0.0 if R is None else np.asarray(R, np.float64)) + gamma * (T64 @ V)
        if fold == FOLD_PI:
            V = (Q * np.asarray(pi, np.float64)).sum(-1)
        elif fold == FOLD_MIN:
            V = Q.min(-1)
        else:
            V = Q.max(-1)
    return Q, V


def episodic_f64(H, T, R, pi=None, max_value=None):
    """finite_horizon.py:11-42.  Returns (Q[H+1,S,A], V[H+1,S]) | None."""
    T, R, pi = _f32(T), _f32(R), _f32(pi)
    S, A, _ = T.shape
    Q = np.zeros((H + 1, S, A), np.float64)
    V = np.zeros((H + 1, S), np.float64)
    rc = lib().orc_episodic_f64(_p(T), _p(R), _p(pi), S, A, H, C.c_double(max_value or 0.0), _p(Q), _p(V))
    return None if rc == OVERFLOW else (Q, V)


def episodic_f32(H, T, R, pi=None):
    T, R, pi = _f32(T), _f32(R), _f32(pi)
    S, A, _ = T.shape
    Q = np.zeros((H + 1, S, A), np.float32)
    V = np.zeros((H + 1, S), np.float32)
    lib().orc_episodic_f32(_p(T), _p(R), _p(pi), S, A, H, _p(Q), _p(V))
    return Q, V


# ---------------------------------------------------------------------------------------------- hardness
def diameter_continuous_f64(T, targets=None, tol=1e-10, max_value=None, max_iter=int(1e6), return_E=False):
    """hardness/measures/diameter.py:76-106 at the fixed point."""
    T = _f32(T)
    S, A, _ = T.shape
    targets = np.arange(S, dtype=np.int32) if targets is None else np.ascontiguousarray(targets, np.int32)
    K = len(targets)
    E = np.zeros((K, S), np.float64) if return_E else None
    d = C.c_double(0)
    sw = C.c_longlong(0)
    rc = lib().orc_diameter_continuous_f64(_p(T), _p(targets), K, S, A, C.c_double(tol), C.c_double(max_value or 0.0),
                                           C.c_longlong(max_iter), _p(E), C.byref(d), C.byref(sw))
    if rc == OVERFLOW:
        return None
    if rc == MAX_ITER:
        raise RuntimeError("max iterations exceeded")
    return (d.value, E, sw.value) if return_E else d.value


def diameter_episodic_f64(T_epi, targets=None, tol=1e-10, max_iter=int(1e6)):
    """hardness/measures/diameter.py:285-318 at the fixed point; T_epi is [H,S,A,S]."""
    T_epi = _f32(T_epi)
    H, S, A, _ = T_epi.shape
    targets = np.arange(S, dtype=np.int32) if targets is None else np.ascontiguousarray(targets, np.int32)
    d = C.c_double(0)
    sw = C.c_longlong(0)
    rc = lib().orc_diameter_episodic_f64(_p(T_epi), _p(targets), len(targets), H, S, A, C.c_double(tol),
                                         C.c_longlong(max_iter), C.byref(d), C.byref(sw))
    if rc == MAX_ITER:
        raise RuntimeError("max iterations exceeded")
    return d.value


def diameter_target_ref_f32(T, es, max_diam=0.0, eps=1e-3):
    T = _f32(T)
    S, A, _ = T.shape
    return float(lib().orc_diameter_target_ref_f32(_p(T), int(es), S, A, C.c_float(max_diam), C.c_float(eps)))


def value_norm_f64(T, V):
    """hardness/measures/value_norm.py:85-87 (Ev indexed by the next state, sic)."""
    T = _f32(T)
    V = np.ascontiguousarray(V, np.float64)
    S, A, _ = T.shape
    return float(lib().orc_value_norm_f64(_p(T), _p(V), S, A))


def gaps_f64(Q, V, mask=None, reg=0.1):
    """hardness/measures/sum_reciprocals_suboptimality_gaps.py:6-28; Q[...,A], V[...] flattened over leading dims."""
    Q = np.ascontiguousarray(Q, np.float64)
    V = np.ascontiguousarray(V, np.float64)
    A = Q.shape[-1]
    NS = V.size
    m = None if mask is None else np.ascontiguousarray(mask, np.uint8).reshape(-1)
    return float(lib().orc_gaps_f64(_p(Q), _p(V), _p(m), C.c_longlong(NS), A, C.c_double(reg)))


def set_threads(n=None):
    """OpenMP threads of the oracle's parallel loops (torchrun exports OMP_NUM_THREADS=1 to its workers, which would
    silently turn the all-core CPU baseline into a single-thread one).  Returns the count in effect."""
    n = int(n or os.cpu_count() or 1)
    lib()  # make sure the OpenMP runtime the oracle links against is loaded
    for name in ("libgomp.so.1", "libgomp.so"):
        try:
            C.CDLL(name).omp_set_num_threads(n)
            return n
        except OSError:
            continue
    return int(os.environ.get("OMP_NUM_THREADS", "1"))


# ---------------------------------------------------------------------------------------------- step
def philox(seed, env, t):
    out = (C.c_uint32 * 4)()
    lib().orc_philox(C.c_uint64(seed), C.c_uint64(env), C.c_uint64(t), out)
    return list(out)


def build_dense_cdf(T, ld=None, f64=False):
    T = _f32(T)
    S, A, _ = T.shape
    ld = ld or ((S + 31) // 32) * 32
    cdf = np.zeros((S, A, ld), np.float64 if f64 else np.float32)
    lib().orc_build_dense_cdf(_p(T), S, A, ld, _p(cdf), int(f64))
    return cdf


class HostTables:
    """numpy-backed tables for orc_env_step; keeps the arrays alive."""

    def __init__(self, S, A, H=0, cdf=None, succ_cum=None, succ_idx=None, succ_len=None, rew_cls_sas=None,
                 rew_cls_sa=None, rew_cls_succ=None, rew_q=None, rmin=0.0, rmax=1.0, start_cum=None, start_idx=None):
        self.keep = dict(
            cdf=None if cdf is None else np.ascontiguousarray(cdf),
            succ_cum=None if succ_cum is None else np.ascontiguousarray(succ_cum, np.float64),
            succ_idx=None if succ_idx is None else np.ascontiguousarray(succ_idx, np.int32),
            succ_len=None if succ_len is None else np.ascontiguousarray(succ_len, np.int32),
            rew_cls_sas=None if rew_cls_sas is None else np.ascontiguousarray(rew_cls_sas, np.uint8),
            rew_cls_sa=None if rew_cls_sa is None else np.ascontiguousarray(rew_cls_sa, np.int32),
            rew_cls_succ=None if rew_cls_succ is None else np.ascontiguousarray(rew_cls_succ, np.int32),
            rew_q=np.ascontiguousarray(rew_q if rew_q is not None else np.zeros((1, 2)), np.float32),
            start_cum=np.ascontiguousarray(start_cum if start_cum is not None else [1.0], np.float64),
            start_idx=np.ascontiguousarray(start_idx if start_idx is not None else [0], np.int32),
        )
        k = self.keep
        t = Tables()
        t.S, t.A, t.H = S, A, H
        t.ld = 0 if k["cdf"] is None else k["cdf"].shape[-1]
        t.Ksucc = 0 if k["succ_cum"] is None else k["succ_cum"].shape[-1]
        for name in ("cdf", "succ_cum", "succ_idx", "succ_len", "rew_cls_sas", "rew_cls_sa", "rew_cls_succ", "rew_q",
                     "start_cum", "start_idx"):
            setattr(t, name, None if k[name] is None else k[name].ctypes.data)
        t.n_cls, t.nq = k["rew_q"].shape
        t.rmin, t.rmax = rmin, rmax
        t.n_start = len(k["start_idx"])
        self.c = t
        self.cdf_is_f64 = k["cdf"] is not None and k["cdf"].dtype == np.float64


def env_reset(tb, N, u_next=None, seed=0, t=0, visits_s=None, env0=0):
    state = np.zeros(N, np.int32)
    h = np.zeros(N, np.int32)
    st = np.zeros(N, np.uint8)
    obs = np.zeros(N, np.int32)
    u = None if u_next is None else np.ascontiguousarray(u_next, np.float64)
    lib().orc_env_reset(C.byref(tb.c), C.c_longlong(N), _p(u), C.c_uint64(seed), C.c_uint64(t), C.c_uint64(env0), _p(state),
                        _p(h), _p(st), _p(obs), _p(visits_s))
    return state, h, st, obs


def env_step(tb, mode, state, h, step_type, action=None, u_next=None, u_rew=None, seed=0, t=0, auto_reset=False,
             visits_s=None, visits_sa=None, env0=0):
    """in-place on state/h/step_type (and action when random). mode: 0 dense f32, 1 dense f64, 2 successor.
    Returns (reward, obs, status)."""
    N = len(state)
    random_actions = action is None
    if random_actions:
        action = np.zeros(N, np.int32)
    assert action.dtype == np.int32 and state.dtype == np.int32 and h.dtype == np.int32 and step_type.dtype == np.uint8
    if u_next is not None:
        u_next = np.ascontiguousarray(u_next, np.float32 if mode == 0 else np.float64)
    if u_rew is not None:
        u_rew = np.ascontiguousarray(u_rew, np.float32)
    reward = np.zeros(N, np.float32)
    obs = np.zeros(N, np.int32)
    rc = lib().orc_env_step(C.byref(tb.c), mode, C.c_longlong(N), _p(action), int(random_actions), _p(u_next),
                            _p(u_rew), C.c_uint64(seed), C.c_uint64(t), C.c_uint64(env0), int(auto_reset), _p(state), _p(h),
                            _p(step_type), _p(reward), _p(obs), _p(visits_s), _p(visits_sa))
    return reward, obs, rc, action


# ---------------------------------------------------------------------------------------------- episodic tensor forms
def episodic_T(H, T, R, start_idx, start_prob):
    """get_episodic_transition_matrix_and_rewards (colosseum/mdp/utils/mdp_creation.py:98-128) restated in numpy.
    Returns (T_epi f32[H,S,A,S], R_epi f32[H,S,A], reach bool[H,S])."""
    T, R = _f32(T), _f32(R)
    S, A, _ = T.shape
    T_epi = np.zeros((H, S, A, S), np.float32)
    for sn, p in zip(start_idx, start_prob):  # :117-120
        T_epi[0, sn] = T[sn]
        T_epi[H - 1, :, :, sn] = p
    for h in range(1, H - 1):  # :121-124
        live = T_epi[h - 1].sum((0, 1)) > 0
        T_epi[h, live] = T[live]
    R_epi = np.tile(R, (H, 1, 1))  # :125-126
    R_epi[-1] = 0.0
    reach = np.zeros((H, S), bool)  # base_finite.py:138-150 (graph reachability == numeric reachability of T)
    reach[0, np.asarray(start_idx)] = True
    adj = (T > 0).any(1)
    for h in range(1, H):
        reach[h] = adj[reach[h - 1]].any(0)
    return T_epi, R_epi, reach


def continuous_form(H, T, R, start_idx, start_prob, nodes):
    """get_continuous_form_episodic_transition_matrix_and_rewards (mdp_creation.py:131-176) restated in numpy;
    `nodes` = the (h, s) pairs in row/column order (the reference's episodic-graph node order)."""
    T, R = _f32(T), _f32(R)
    S, A, _ = T.shape
    nodes = [(int(h), int(s)) for h, s in nodes]
    n = len(nodes)
    index = {hs: i for i, hs in enumerate(nodes)}
    T_cf = np.zeros((n, A, n), np.float32)
    R_cf = np.zeros((n, A), np.float32)
    for i, (h, s) in enumerate(nodes):
        R_cf[i] = R[s]
        if h == H - 1:  # :166-168 -- sic: the column is node_to_index[sn], the start state's index in the ORIGINAL
            # MDP, not the position of (0, sn) in the node list; kept (the reference's episodic value norm runs on it)
            for sn, p in zip(start_idx, start_prob):
                T_cf[i, :, int(sn)] = p
        else:  # :170-172 (graph successors == next states with positive probability under some action)
            for ns in np.nonzero((T[s] > 0).any(0))[0]:
                T_cf[i, :, index[(h + 1, int(ns))]] = T[s, :, ns]
    return T_cf, R_cf


# ---------------------------------------------------------------------------------------------- extended VI (UCRL2)
def extended_vi_f32(T, est, beta_r, beta_p, r_max, eps=1e-3, max_iter=int(1e6)):
    """infinite_horizon.py:67-118 + _max_proba :222-251, restated loop for loop in C.  Returns (span, Q, V, iters) | None."""
    T, est = _f32(T), _f32(est)
    S, A, _ = T.shape
    br = np.ascontiguousarray(beta_r, np.float64).reshape(S, A)
    bp = np.asarray(beta_p, np.float64)
    bp = np.ascontiguousarray(bp[..., 0] if bp.ndim == 3 else bp).reshape(S, A)
    Q = np.zeros((S, A), np.float32)
    V = np.zeros(S, np.float32)
    span = C.c_double(0)
    it = C.c_longlong(0)
    rc = lib().orc_extended_vi_f32(_p(T), _p(est), _p(br), _p(bp), S, A, C.c_double(r_max), C.c_double(eps),
                                   C.c_longlong(max_iter), _p(Q), _p(V), C.byref(span), C.byref(it))
    return None if rc == MAX_ITER else (span.value, Q, V, it.value)


# ---------------------------------------------------------------------------------------------- policy Markov chain
def policy_chain(T, R, pi):
    """markov_chain.py:34-51: (min(1, einsum('saj,sa->sj', T, pi)), einsum('sa,sa->s', R, pi))"""
    T, R, pi = _f32(T), _f32(R), _f32(pi)
    return np.minimum(1.0, np.einsum("saj,sa->sj", T, pi)), np.einsum("sa,sa->s", R, pi)


def stationary_distribution_f64(P, x0, tol=1e-14, max_iter=int(1e7)):
    """limit of x0 under the lazy chain (I + P)/2 in fp64 (repeated squaring of the matrix: log2(n) products) --
    for a chain with one recurrent class this is THE stationary distribution of markov_chain.py:64-137"""
    M = 0.5 * (np.eye(len(P)) + np.asarray(P, np.float64))
    M = M / M.sum(-1, keepdims=True)
    for _ in range(200):
        M2 = M @ M
        M2 = M2 / M2.sum(-1, keepdims=True)  # rows of a float32 T sum to 1 +- 1e-7: keep the powers stochastic
        if np.abs(M2 - M).max() < tol:
            M = M2
            break
        M = M2
    x = np.asarray(x0, np.float64) @ M
    return x / x.sum()


# ---------------------------------------------------------------------------------------------- agent loops
class _ActorArgs(C.Structure):
    _fields_ = [("epsilon_schedule", C.c_void_p), ("temperature_schedule", C.c_void_p), ("t0", C.c_longlong),
                ("len", C.c_int), ("boltzmann", C.c_int), ("boltzmann_temperature", C.c_double)]


def set_actor(args, keep, epsilon_greedy, boltzmann_temperature, t0, n):
    """fills args.epsilon_greedy and args.actor for the interaction counts t0 .. t0 + n - 1: constants as they are,
    functions of the actor's interaction counter (Q_values_actor.py:58-78) tabulated.  `keep` holds the tables alive."""
    a = args.actor
    a.t0, a.len = int(t0), int(n)
    a.epsilon_schedule = a.temperature_schedule = None
    args.epsilon_greedy = -1.0
    if callable(epsilon_greedy):
        keep["eps"] = np.array([float(epsilon_greedy(t0 + k)) for k in range(n)], np.float64)
        a.epsilon_schedule = _p(keep["eps"])
        args.epsilon_greedy = 0.0
    elif epsilon_greedy is not None:
        args.epsilon_greedy = float(epsilon_greedy)
    a.boltzmann = int(boltzmann_temperature is not None)
    if callable(boltzmann_temperature):
        keep["temp"] = np.array([float(boltzmann_temperature(t0 + k)) for k in range(n)], np.float64)
        a.temperature_schedule = _p(keep["temp"])
    elif boltzmann_temperature is not None:
        a.boltzmann_temperature = float(boltzmann_temperature)


def boltzmann_action(q, temperature, u):
    """Q_values_actor.py:73-78 given the uniform numpy's choice consumes: index of the action"""
    q = _f32(q)
    return int(lib().orc_boltzmann_action(_p(q), len(q), C.c_double(temperature), C.c_double(u)))


class _QLArgs(C.Structure):
    _fields_ = [
        ("N", C.c_longlong), ("seed", C.c_uint64), ("env0", C.c_uint64),
        ("state", C.c_void_p), ("h", C.c_void_p), ("cnt", C.c_void_p), ("Q", C.c_void_p), ("Q_main", C.c_void_p),
        ("V", C.c_void_p), ("mu", C.c_void_p), ("sigma", C.c_void_p), ("beta", C.c_void_p),
        ("ucb_type", C.c_int),
        ("c_1", C.c_double), ("c_2", C.c_double), ("min_at", C.c_double), ("log_term", C.c_double),
        ("sqrt_h7sa", C.c_double), ("H_eff", C.c_double), ("gamma", C.c_double), ("span_approx", C.c_double),
        ("epsilon_greedy", C.c_double), ("cum_reward", C.c_void_p), ("n_episodes", C.c_void_p),
        ("trace", C.c_void_p), ("actor", _ActorArgs),
    ]


class QLearningLoops:
    """CPU restatement of N Q-learning agent/MDP loops (orc_qlearning_steps in colo_oracle.c; reference:
    agent/agents/episodic/q_learning.py:28-103, agent/agents/infinite_horizon/q_learning.py:48-111,
    agent/actors/Q_values_actor.py:58-82, experiment/agent_mdp_interaction.py:238-298).  `tb` is a HostTables with
    successor tables."""

    def __init__(self, tb, n_loops, optimization_horizon, seed=0, env0=0, epsilon_greedy=None, *, p=0.05, c_1=1.0,
                 c_2=None, min_at=0.0, UCB_type="hoeffding", confidence=0.95, span_approx_weight=1.0, h_weight=1.0,
                 boltzmann_temperature=None):
        self.tb, self.N, self.seed, self.env0 = tb, int(n_loops), int(seed), int(env0)
        self._explore, self._keep = (epsilon_greedy, boltzmann_temperature), {}
        S, A, H = tb.c.S, tb.c.A, tb.c.H
        self.episodic = H > 0
        N = self.N
        self.state, self.h, _, _ = env_reset(tb, N, seed=seed, t=0, env0=env0)
        self.t = 1
        self.cum_reward = np.zeros(N, np.float64)
        self.n_episodes = np.zeros(N, np.int64)
        a = _QLArgs()
        a.N, a.seed, a.env0 = N, self.seed, self.env0
        if self.episodic:
            self.cnt = np.ones((N, H, S, A), np.int32)
            self.Q = np.full((N, H, S, A), H, np.float32)
            self.V = np.zeros((N, H + 1, S), np.float32)
            a.ucb_type = 0
            if UCB_type.lower() == "bernstein":
                self.mu, self.sigma, self.beta = (np.zeros((N, H, S, A), np.float32) for _ in range(3))
                a.mu, a.sigma, a.beta, a.ucb_type, a.c_2 = _p(self.mu), _p(self.sigma), _p(self.beta), 1, float(c_2)
            a.c_1, a.min_at = float(c_1), float(min_at)
            a.log_term = float(np.log(S * A * optimization_horizon / p))
            a.sqrt_h7sa = float(np.sqrt(H ** 7 * S * A))
        else:
            T_ = optimization_horizon
            span = span_approx_weight
            Hf = float(h_weight * min(np.sqrt(span * T_ / S / A), (T_ / S / A / np.log(4 * T_ / confidence)) ** 0.333))
            self.H_eff, self.gamma = Hf, 1 - 1 / Hf
            self.cnt = np.zeros((N, S, A), np.int32)
            self.Q = np.full((N, S, A), Hf, np.float32)
            self.Q_main = np.full((N, S, A), Hf, np.float32)
            self.V = np.full((N, S), Hf, np.float32)
            a.Q_main = _p(self.Q_main)
            a.min_at = float(min_at if min_at > 0.009 else 0)
            a.H_eff, a.gamma, a.span_approx = Hf, float(self.gamma), float(span)
            a.log_term = float(np.log(2 * T_ / confidence))
        a.state, a.h, a.cnt, a.Q, a.V = _p(self.state), _p(self.h), _p(self.cnt), _p(self.Q), _p(self.V)
        a.cum_reward, a.n_episodes = _p(self.cum_reward), _p(self.n_episodes)
        self.args = a

    def steps(self, n_steps, trace=False):
        tr = np.zeros((n_steps, self.N, 4), np.int32) if trace else None
        self.args.trace = _p(tr)
        set_actor(self.args, self._keep, *self._explore, self.t, n_steps)
        rc = lib().orc_qlearning_steps(C.byref(self.tb.c), C.byref(self.args), int(self.episodic), int(n_steps),
                                       C.c_uint64(self.t))
        assert rc == 0
        self.t += n_steps
        return tr


class _PsrlArgs(C.Structure):
    _fields_ = [
        ("N", C.c_longlong), ("seed", C.c_uint64), ("env0", C.c_uint64),
        ("state", C.c_void_p), ("h", C.c_void_p), ("Q", C.c_void_p), ("dir_hyper", C.c_void_p),
        ("nig_hyper", C.c_void_p), ("epsilon_greedy", C.c_double), ("cum_reward", C.c_void_p),
        ("n_episodes", C.c_void_p), ("trace", C.c_void_p), ("reward_model", C.c_int),
        ("actor", _ActorArgs),
    ]


class PSRLLoops:
    """CPU restatement of N PSRLEpisodic loops BETWEEN posterior samples (orc_psrl_steps): the caller supplies the Q
    of each loop's sampled model (`set_q`), as the product does after colo_sample_* + episodic VI.  Priors as
    BayesianMDPModel (agent/mdp_models/bayesian_model.py:44-57) and N_NIG.__init__'s reparametrisation
    (bayesian_models/conjugate_rewards.py:45-54)."""

    def __init__(self, tb, n_loops, seed=0, env0=0, epsilon_greedy=None, rewards_prior_prms=None,
                 transitions_prior_prms=None, reward_prior_model="N_NIG", boltzmann_temperature=None):
        self.tb, self.N, self.seed, self.env0 = tb, int(n_loops), int(seed), int(env0)
        self._explore, self._keep = (epsilon_greedy, boltzmann_temperature), {}
        self.reward_model = {"N_NIG": 0, "N_N": 1}[reward_prior_model]
        S, A, H = tb.c.S, tb.c.A, tb.c.H
        N = self.N
        self.state, self.h, _, _ = env_reset(tb, N, seed=seed, t=0, env0=env0)
        self.t = 1
        self.cum_reward = np.zeros(N, np.float64)
        self.n_episodes = np.zeros(N, np.int64)
        rp = [tb.c.rmax, 1, 1, 1] if rewards_prior_prms is None else rewards_prior_prms
        tp = [1.0 / S] if transitions_prior_prms is None else transitions_prior_prms
        if self.reward_model == 1:  # N_N: (mu, tau), stored in the first two of the four slots
            hp = np.zeros((S, A, 4), np.float32)
            hp[..., :2] = np.tile(rp, (S, A, 1)).astype(np.float32)
        else:
            hp = np.tile(rp, (S, A, 1)).astype(np.float32)
            mu, n_mu, tau, n_tau = hp[..., 0].copy(), hp[..., 1].copy(), hp[..., 2].copy(), hp[..., 3].copy()
            hp[..., 0], hp[..., 1], hp[..., 2], hp[..., 3] = mu, n_mu, n_tau * 0.5, (0.5 * n_tau) / tau
        self.nig_hyper = np.tile(hp, (N, 1, 1, 1)).astype(np.float32)
        self.dir_hyper = np.tile(np.float32(tp[0]), (N, S, A, S)).astype(np.float32)
        self.Q = np.zeros((N, H + 1, S, A), np.float32)
        a = _PsrlArgs()
        a.N, a.seed, a.env0 = N, self.seed, self.env0
        a.state, a.h, a.Q = _p(self.state), _p(self.h), _p(self.Q)
        a.dir_hyper, a.nig_hyper = _p(self.dir_hyper), _p(self.nig_hyper)
        a.cum_reward, a.n_episodes = _p(self.cum_reward), _p(self.n_episodes)
        a.reward_model = self.reward_model
        self.args = a

    def set_q(self, Q):
        self.Q[...] = Q

    def steps(self, n_steps, trace=False):
        tr = np.zeros((n_steps, self.N, 4), np.int32) if trace else None
        self.args.trace = _p(tr)
        set_actor(self.args, self._keep, *self._explore, self.t, n_steps)
        rc = lib().orc_psrl_steps(C.byref(self.tb.c), C.byref(self.args), int(n_steps), C.c_uint64(self.t))
        assert rc == 0
        self.t += n_steps
        return tr


# ------------------------------------------------------------------------------------------------ UCRL2Continuous
class _Ucrl2Args(C.Structure):
    _fields_ = [
        ("N", C.c_longlong), ("seed", C.c_uint64), ("env0", C.c_uint64),
        ("state", C.c_void_p), ("t", C.c_void_p), ("cum_reward", C.c_void_p), ("Q", C.c_void_p),
        ("Nsas", C.c_void_p), ("Nsa", C.c_void_p), ("P", C.c_void_p), ("est_r", C.c_void_p), ("var_r", C.c_void_p),
        ("hold", C.c_void_p), ("nu", C.c_void_p), ("seen", C.c_void_p), ("ep_len", C.c_void_p), ("ep_log", C.c_void_p),
        ("log_cap", C.c_int), ("ended", C.c_void_p), ("iteration", C.c_void_p), ("episode", C.c_void_p),
        ("delta", C.c_void_p), ("epsilon_greedy", C.c_double), ("trace", C.c_void_p), ("trace_t0", C.c_longlong),
        ("trace_steps", C.c_int),
        ("actor", _ActorArgs),
    ]


class UCRL2Loops:
    """CPU restatement of N UCRL2Continuous loops (colosseum/agent/agents/infinite_horizon/ucrl2.py:34-357): the three
    C functions orc_ucrl2_{steps,bounds,model_update} plus orc_extended_vi_f32 as the planner, driven in the rounds the
    product uses (run every loop to its episode end or the target time; episode_end_update for the loops that wait).
    `planner(loop, episode, T, est_r, beta_r, beta_p, r_max)` -> (span, Q, V) may replace the planner and record=True
    keeps every (loop, episode) -> (Q, beta_r, beta_p): the GPU tests feed both sides the same Q so that trajectories
    and model tables compare bit for bit."""

    def __init__(self, tb, n_loops, optimization_horizon, seed=0, env0=0, alpha_r=1.0, alpha_p=1.0,
                 bound_type_p="_chernoff", bound_type_rew="_chernoff", epsilon_greedy=None, planner=None, record=False,
                 boltzmann_temperature=None):
        assert bound_type_p in ("_chernoff", "bernstein") and bound_type_rew == "_chernoff"
        self._explore, self._keep = (epsilon_greedy, boltzmann_temperature), {}
        self.history = {} if record else None
        assert tb.c.H == 0, "UCRL2Continuous needs a continuous MDP"
        self.tb, self.N, self.seed, self.env0 = tb, int(n_loops), int(seed), int(env0)
        S, A, N = tb.c.S, tb.c.A, self.N
        self.S, self.A = S, A
        self.alpha_r, self.alpha_p, self.bernstein_p = float(alpha_r), float(alpha_p), int(bound_type_p == "bernstein")
        self.r_max = float(tb.c.rmax)
        self.planner = planner
        self.state, _, _, _ = env_reset(tb, N, seed=seed, t=0, env0=env0)
        self.t = np.ones(N, np.int64)  # counter 0 was the reset draw
        self.cum_reward = np.zeros(N, np.float64)
        self.Q = np.zeros((N, S, A), np.float32)
        self.V = np.zeros((N, S), np.float32)
        self.Nsas = np.zeros((N, S, A, S), np.int32)
        self.Nsa = np.zeros((N, S, A), np.int32)
        self.P = (np.ones((N, S, A, S), np.float32) / S).astype(np.float32)              # ucrl2.py:148
        self.est_r = (np.ones((N, S, A), np.float32) * tb.c.rmax).astype(np.float32)     # :150-152
        self.var_r = np.zeros((N, S, A), np.float32)
        self.hold = np.ones((N, S, A), np.float32)
        self.nu = np.zeros((N, S, A), np.int32)
        self.seen = np.zeros((N, S, A), np.int32)
        self.log_cap = (int(optimization_horizon) + S * A) // 2 + 2
        self.ep_len = np.zeros(N, np.int32)
        self.ep_log = np.zeros((N, self.log_cap, 2), np.int32)
        self.ended = np.zeros(N, np.int32)
        self.iteration = np.zeros(N, np.int64)
        self.episode = np.zeros(N, np.int64)
        self.delta = np.ones(N, np.float64)
        self.span_value = np.zeros(N, np.float64)
        self.episode_ends = [[] for _ in range(N)]  # interaction times at which each loop's episodes ended
        a = _Ucrl2Args()
        a.N, a.seed, a.env0 = N, self.seed, self.env0
        for k in ("state", "t", "cum_reward", "Q", "Nsas", "Nsa", "P", "est_r", "var_r", "hold", "nu", "seen", "ep_len",
                  "ep_log", "ended", "iteration", "episode", "delta"):
            setattr(a, k, _p(getattr(self, k)))
        a.log_cap = self.log_cap
        self.args = a
        self.episode_end_update(np.arange(N, dtype=np.int32))  # before_start_interacting (:194-195)

    def bounds(self, idx):
        idx = np.ascontiguousarray(idx, np.int32)
        m = len(idx)
        br = np.zeros((m, self.S, self.A), np.float64)
        bp = np.zeros((m, self.S, self.A), np.float64)
        rc = lib().orc_ucrl2_bounds(C.byref(self.args), self.S, self.A, _p(idx), m, C.c_double(self.alpha_r),
                                    C.c_double(self.alpha_p), C.c_double(self.r_max), self.bernstein_p, _p(br), _p(bp))
        assert rc == 0
        return br, bp

    def episode_end_update(self, idx):
        """ucrl2.py:183-192 for the loops in idx: bounds -> extended VI -> model_update (in the reference's order)"""
        idx = np.ascontiguousarray(idx, np.int32)
        had_data = self.ep_len[idx] > 0
        br, bp = self.bounds(idx)
        for k, i in enumerate(idx):
            if self.planner is not None:
                res = self.planner(int(i), int(self.episode[i]), self.P[i], self.est_r[i], br[k], bp[k], self.r_max)
            else:
                res = extended_vi_f32(self.P[i], self.est_r[i], br[k], bp[k], self.r_max)
            if res is not None:
                span, Q, V = res[:3]
                self.Q[i], self.V[i] = Q, V
                self.span_value[i] = span / self.r_max
            if self.history is not None:  # (loop, episode number) -> what the planner saw and returned
                self.history[(int(i), int(self.episode[i]))] = (self.Q[i].copy(), br[k].copy(), bp[k].copy())
        upd = np.ascontiguousarray(idx[had_data], np.int32)
        rc = lib().orc_ucrl2_model_update(C.byref(self.args), self.S, self.A, _p(upd), len(upd))
        assert rc == 0
        self.ended[idx] = 0

    def steps(self, n_steps, trace=False):
        """n_steps interactions for every loop.  trace: i32 [n_steps, N, 4] = (s_t, a_t, s_tp1, reward bits)"""
        t0 = int(self.t[0])
        assert (self.t == t0).all()
        tr = np.zeros((n_steps, self.N, 4), np.int32) if trace else None
        self.args.trace, self.args.trace_t0, self.args.trace_steps = _p(tr), t0, int(n_steps)
        target = t0 + int(n_steps)
        set_actor(self.args, self._keep, *self._explore, t0, n_steps)
        while True:
            rc = lib().orc_ucrl2_steps(C.byref(self.tb.c), C.byref(self.args), C.c_longlong(target))
            assert rc == 0
            assert (self.ended != 2).all(), "episode log overflow: optimization_horizon is too small for this run"
            idx = np.nonzero(self.ended)[0].astype(np.int32)
            if len(idx) == 0:
                break
            for i in idx:
                self.episode_ends[i].append(int(self.t[i]))
            self.episode_end_update(idx)
        self.args.trace = None
        return tr


# ------------------------------------------------------------------------------------------------- PSRLContinuous
class _PsrlcArgs(C.Structure):
    _fields_ = [
        ("N", C.c_longlong), ("seed", C.c_uint64), ("env0", C.c_uint64),
        ("state", C.c_void_p), ("t", C.c_void_p), ("cum_reward", C.c_void_p), ("Q", C.c_void_p), ("psi", C.c_int),
        ("dir_hyper", C.c_void_p), ("nig_hyper", C.c_void_p), ("reward_model", C.c_int),
        ("Nsas", C.c_void_p), ("Nsa", C.c_void_p), ("nu", C.c_void_p), ("ended", C.c_void_p), ("episode", C.c_void_p),
        ("epsilon_greedy", C.c_double), ("trace", C.c_void_p), ("trace_t0", C.c_longlong), ("trace_steps", C.c_int),
        ("actor", _ActorArgs),
    ]


def psrlc_parameters(S, A, optimization_horizon, p=0.05, psi_weight=1.0, omega_weight=1.0, kappa_weight=1.0,
                     eta_weight=1.0, max_psi=60, no_optimistic_sampling=False):
    """psi, omega, kappa, eta of PSRLContinuous.__init__ (posterior_sampling.py:20-115, :268-303), in its expressions"""
    T = optimization_horizon
    no_opt = bool(no_optimistic_sampling or (S ** 2 * A) > 6_000_000)
    psi = min(max_psi, max(2, int(psi_weight * (S * np.log(S * A / p)))))
    omega = omega_weight * np.log(T / p)
    kappa = kappa_weight * np.log(T / p)
    eta = max(5, min(10 * S, eta_weight * (np.sqrt(T * S / A) + 12 * omega * S ** 4)))
    return dict(psi=1 if no_opt else psi, omega=omega, kappa=kappa, eta=0.0 if no_opt else float(eta),
                no_optimistic_sampling=no_opt)


def psrlc_simple_rows(Nsas, z):
    """the "simple sampling" rows of optimistic_sampling (posterior_sampling.py:424-446) for ONE sample: P_minus with the
    missing mass on state z, float32 [S,A,S].  The reference's own numpy expressions."""
    S = Nsas.shape[-1]
    Nsum = Nsas.sum(-1)
    P_hat = Nsas / np.maximum(Nsum[..., None], 1)
    N = np.maximum(Nsas, 1)
    P_minus = P_hat - np.minimum(np.sqrt(3 * P_hat * np.log(4 * S) / N) + 3 * np.log(4 * S) / N, P_hat)
    summing = 1 - P_minus.sum(-1)
    P_minus[:, :, z] += summing
    return P_minus.astype(np.float32)


def psrlc_z(seed, env, episode, q, S):
    """the state that receives the missing mass in sample q of the loop's `episode`-th re-planning"""
    w = philox(seed ^ 0xC2B2AE3D27D4EB4F, env, episode * 64 + q)
    return int((int(w[0]) * S) >> 32)


class PSRLCLoops:
    """CPU restatement of N PSRLContinuous loops BETWEEN re-plannings (orc_psrlc_steps): the caller supplies the
    extended q-values of each loop (`Q[i] = ...` inside `planner(loops, idx)`), as the product does after
    colo_psrlc_sample_models + the discounted VI.  Priors as BayesianMDPModel (bayesian_model.py:44-57)."""

    def __init__(self, tb, n_loops, psi, seed=0, env0=0, epsilon_greedy=None, rewards_prior_prms=None,
                 transitions_prior_prms=None, reward_prior_model="N_NIG", planner=None, boltzmann_temperature=None):
        assert tb.c.H == 0
        self._explore, self._keep = (epsilon_greedy, boltzmann_temperature), {}
        self.tb, self.N, self.seed, self.env0, self.psi = tb, int(n_loops), int(seed), int(env0), int(psi)
        self.reward_model = {"N_NIG": 0, "N_N": 1}[reward_prior_model]
        S, A, N = tb.c.S, tb.c.A, self.N
        self.S, self.A = S, A
        self.planner = planner
        self.state, _, _, _ = env_reset(tb, N, seed=seed, t=0, env0=env0)
        self.t = np.ones(N, np.int64)
        self.cum_reward = np.zeros(N, np.float64)
        rp = [tb.c.rmax, 1, 1, 1] if rewards_prior_prms is None else rewards_prior_prms
        tp = [1.0 / S] if transitions_prior_prms is None else transitions_prior_prms
        if self.reward_model == 1:
            hp = np.zeros((S, A, 4), np.float32)
            hp[..., :2] = np.tile(rp, (S, A, 1)).astype(np.float32)
        else:
            hp = np.tile(rp, (S, A, 1)).astype(np.float32)
            mu, n_mu, tau, n_tau = hp[..., 0].copy(), hp[..., 1].copy(), hp[..., 2].copy(), hp[..., 3].copy()
            hp[..., 0], hp[..., 1], hp[..., 2], hp[..., 3] = mu, n_mu, n_tau * 0.5, (0.5 * n_tau) / tau
        self.nig_hyper = np.tile(hp, (N, 1, 1, 1)).astype(np.float32)
        self.dir_hyper = np.tile(np.float32(tp[0]), (N, S, A, S)).astype(np.float32)
        self.Q = np.zeros((N, S, A * self.psi), np.float32)
        self.Nsas = np.zeros((N, S, A, S), np.int32)
        self.Nsa = np.zeros((N, S, A), np.int32)
        self.nu = np.zeros((N, S, A), np.int32)
        self.ended = np.zeros(N, np.int32)
        self.episode = np.zeros(N, np.int64)
        self.episode_ends = [[] for _ in range(N)]
        a = _PsrlcArgs()
        a.N, a.seed, a.env0, a.psi, a.reward_model = N, self.seed, self.env0, self.psi, self.reward_model
        for k in ("state", "t", "cum_reward", "Q", "dir_hyper", "nig_hyper", "Nsas", "Nsa", "nu", "ended", "episode"):
            setattr(a, k, _p(getattr(self, k)))
        self.args = a
        self.episode_end_update(np.arange(N))  # before_start_interacting (posterior_sampling.py:378-382)

    def episode_end_update(self, idx):
        self.planner(self, np.asarray(idx))   # fills self.Q[idx]
        self.nu[idx] = 0
        self.ended[idx] = 0
        self.episode[idx] += 1

    def steps(self, n_steps, trace=False):
        t0 = int(self.t[0])
        assert (self.t == t0).all()
        tr = np.zeros((n_steps, self.N, 4), np.int32) if trace else None
        self.args.trace, self.args.trace_t0, self.args.trace_steps = _p(tr), t0, int(n_steps)
        target = t0 + int(n_steps)
        set_actor(self.args, self._keep, *self._explore, t0, n_steps)
        while True:
            rc = lib().orc_psrlc_steps(C.byref(self.tb.c), C.byref(self.args), C.c_longlong(target))
            assert rc == 0
            idx = np.nonzero(self.ended)[0]
            if len(idx) == 0:
                break
            for i in idx:
                self.episode_ends[i].append(int(self.t[i]))
            self.episode_end_update(idx)
        self.args.trace = None
        return tr
