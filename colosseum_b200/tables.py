"""Host-side table extraction: from a reference `BaseMDP` (duck-typed, the package is never imported here), from
raw successor lists, or from a dense T, to the flat arrays the step kernels read (`colo_mdp_tables`).

What is extracted and where it comes from in the reference (paths under /root/reference/colosseum/):
  successor lists / probabilities in sampler order   mdp/utils/mdp_creation.py:276-310 (NextStateSampler per (s,a))
  start distribution in sampler order                mdp/base.py:463-503  (_starting_node_sampler)
  reward distribution per (s, a, s')                 mdp/base.py:1170-1185 (get_reward_distribution)
  rewards_range, H                                   mdp/base.py, mdp/base_finite.py:33-122
The arrays are plain numpy (so this module imports and is tested without a GPU); `.to_device()` uploads them.
"""
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np

DEFAULT_NQ = 4097  # knots of the tabulated reward quantile functions (uniform grid on [0,1])


def running_sum(p):
    """itertools.accumulate(p) as CPython's random.choices does it: sequential fp64 adds along the last axis."""
    return np.cumsum(np.asarray(p, np.float64), axis=-1)


def quantile_table(kind: str, args: Tuple[float, ...], nq: int = DEFAULT_NQ) -> np.ndarray:
    """Quantile function of one reward distribution on a uniform grid of nq knots.
    kinds follow colosseum/utils/miscellanea.py:253-270 `get_dist`: 'deterministic' or any scipy.stats name."""
    grid = np.linspace(0.0, 1.0, nq)
    if kind == "deterministic":
        return np.full(nq, float(args[0]), np.float32)
    import scipy.stats

    dist = getattr(scipy.stats, kind)(*args)
    q = dist.ppf(grid)
    lo, hi = dist.support()
    q[0] = lo if np.isfinite(lo) else dist.ppf(0.5 / nq)
    q[-1] = hi if np.isfinite(hi) else dist.ppf(1.0 - 0.5 / nq)
    return q.astype(np.float32)


@dataclass
class MDPTables:
    S: int
    A: int
    H: int = 0  # 0 = continuous
    rmin: float = 0.0
    rmax: float = 1.0
    start_idx: np.ndarray = None  # i32 [n_start]
    start_cum: np.ndarray = None  # f64 [n_start]
    # successor form
    succ_idx: Optional[np.ndarray] = None  # i32 [S,A,K] padded with the last real successor
    succ_cum: Optional[np.ndarray] = None  # f64 [S,A,K] padded with +inf
    succ_len: Optional[np.ndarray] = None  # i32 [S,A]
    rew_cls_succ: Optional[np.ndarray] = None  # i32 [S,A,K]
    # dense form
    T: Optional[np.ndarray] = None  # f32 [S,A,S]
    rew_cls_sas: Optional[np.ndarray] = None  # u8 [S,A,S]
    rew_cls_sa: Optional[np.ndarray] = None  # i32 [S,A]
    # rewards
    rew_kinds: List[Tuple[str, Tuple[float, ...]]] = field(default_factory=lambda: [("deterministic", (0.0,))])
    rew_q: np.ndarray = None  # f32 [n_cls, nq]
    # reference attribute surface carried along for the drop-in (mdp/base.py:463-503): expected rewards as the
    # reference built them, and the node <-> index maps (None when the tables do not come from an MDP object)
    R: Optional[np.ndarray] = None  # f32 [S,A]
    node_to_index: Optional[dict] = None
    index_to_node: Optional[dict] = None

    # ------------------------------------------------------------------------------------------ builders
    @staticmethod
    def _finish_rewards(tb, nq):
        tb.rew_q = np.stack([quantile_table(k, tuple(a), nq) for k, a in tb.rew_kinds]).astype(np.float32)
        return tb

    @classmethod
    def from_successors(cls, S, A, succ_idx, succ_prob, succ_len, rew_cls, rew_kinds, start_idx, start_prob, H=0,
                        rewards_range=(0.0, 1.0), T=None, nq=DEFAULT_NQ):
        """succ_* are [S,A,K] in the samplers' own order (padding beyond succ_len is ignored)."""
        succ_idx = np.asarray(succ_idx, np.int32).copy()
        succ_len = np.asarray(succ_len, np.int32)
        K = succ_idx.shape[-1]
        k = np.arange(K)[None, None, :]
        valid = k < succ_len[..., None]
        cum = running_sum(np.where(valid, succ_prob, 0.0))
        cum = np.where(valid, cum, np.inf)
        last = np.take_along_axis(succ_idx, (succ_len - 1)[..., None].astype(np.int64), -1)
        succ_idx = np.where(valid, succ_idx, last).astype(np.int32)
        rc = np.asarray(rew_cls, np.int32)
        last_c = np.take_along_axis(rc, (succ_len - 1)[..., None].astype(np.int64), -1)
        rc = np.where(valid, rc, last_c).astype(np.int32)
        tb = cls(S=S, A=A, H=int(H), rmin=float(rewards_range[0]), rmax=float(rewards_range[1]),
                 start_idx=np.asarray(start_idx, np.int32), start_cum=running_sum(start_prob),
                 succ_idx=succ_idx, succ_cum=cum, succ_len=succ_len, rew_cls_succ=rc,
                 rew_kinds=[(k_, tuple(a_)) for k_, a_ in rew_kinds])
        assert len(tb.rew_kinds) <= 256, "at most 256 distinct reward distributions per MDP"
        # dense twins: T (duplicates summed, as mdp/utils/mdp_creation.py:80 does) and the (s,a,s') class table
        if T is None:
            # the reference's own accumulation: a float32 array, `T[s,a,s'] += p` once per successor in sampler order
            # (mdp_creation.py:67-80; under numpy 2 the Python float p is rounded to float32 before the add).
            # np.add.at is unbuffered and applies the updates in index order, i.e. in that same order.
            T = np.zeros((S, A, S), np.float32)
            sa = np.broadcast_to(np.arange(S * A).reshape(S, A, 1), succ_idx.shape)
            np.add.at(T.reshape(S * A, S), (sa[valid], succ_idx[valid]),
                      np.asarray(succ_prob, np.float64)[valid].astype(np.float32))
        tb.T = np.ascontiguousarray(T, np.float32)
        sas = np.zeros((S, A, S), np.uint8)
        sa = np.broadcast_to(np.arange(S * A).reshape(S, A, 1), succ_idx.shape)
        sas.reshape(S * A, S)[sa[valid], succ_idx[valid]] = rc[valid]
        tb.rew_cls_sas = sas
        return cls._finish_rewards(tb, nq)

    @classmethod
    def from_dense(cls, T, rew_kinds=None, rew_cls_sa=None, rew_cls_sas=None, start_idx=(0,), start_prob=(1.0,), H=0,
                   rewards_range=(0.0, 1.0), nq=DEFAULT_NQ):
        """Any dense T[S,A,S] (synthetic MDPs, CustomMDP: successors = non-zeros of the row in index order,
        colosseum/mdp/custom_mdp.py:82-89)."""
        T = np.ascontiguousarray(T, np.float32)
        S, A, _ = T.shape
        tb = cls(S=S, A=A, H=int(H), rmin=float(rewards_range[0]), rmax=float(rewards_range[1]),
                 start_idx=np.asarray(start_idx, np.int32), start_cum=running_sum(start_prob), T=T,
                 rew_kinds=[(k_, tuple(a_)) for k_, a_ in (rew_kinds or [("deterministic", (0.0,))])])
        if rew_cls_sas is not None:
            tb.rew_cls_sas = np.ascontiguousarray(rew_cls_sas, np.uint8)
        if rew_cls_sa is not None:
            tb.rew_cls_sa = np.ascontiguousarray(rew_cls_sa, np.int32)
        return cls._finish_rewards(tb, nq)

    @classmethod
    def from_golden(cls, g, nq=DEFAULT_NQ):
        """tests/golden/inst_*.npz (arrays recorded from the reference by tests/golden/make_golden.py)."""
        import json

        kinds = [(k, tuple(a)) for k, a in json.loads(str(g["rew_kinds"]))]
        tb = cls.from_successors(int(g["n_states"]), int(g["n_actions"]), g["succ_idx"], g["succ_prob"],
                                 g["succ_len"], g["rew_cls"], kinds, g["start_idx"], g["start_prob"],
                                 H=int(g["H"]), rewards_range=tuple(g["rewards_range"]), T=g["T"], nq=nq)
        if "R" in getattr(g, "files", g):
            tb.R = np.asarray(g["R"], np.float32)
        return tb

    @classmethod
    def from_mdp(cls, mdp, nq=DEFAULT_NQ):
        """A reference `BaseMDP` instance (colosseum/mdp/base.py:45).  Uses only its public/duck-typed surface:
        G.nodes order == state index (base.py:488-492), get_info_class(node).transition_distributions[a]
        (.next_nodes/.probs), _starting_node_sampler, get_reward_distribution, rewards_range, H, T."""
        nodes = list(mdp.G.nodes)
        n2i = mdp.node_to_index
        S, A = mdp.n_states, mdp.n_actions
        tds = [[mdp.get_info_class(n).transition_distributions[a] for a in range(A)] for n in nodes]
        K = max(len(td.next_nodes) for row in tds for td in row)
        succ_idx = np.zeros((S, A, K), np.int32)
        succ_prob = np.zeros((S, A, K), np.float64)
        succ_len = np.zeros((S, A), np.int32)
        rew_cls = np.zeros((S, A, K), np.int32)
        kinds: List[Tuple[str, Tuple[float, ...]]] = []
        for i, n in enumerate(nodes):
            for a in range(A):
                td = tds[i][a]
                succ_len[i, a] = len(td.next_nodes)
                for k, (nn, p) in enumerate(zip(td.next_nodes, td.probs)):
                    succ_idx[i, a, k] = n2i[nn]
                    succ_prob[i, a, k] = p
                    d = mdp.get_reward_distribution(n, a, nn)
                    key = (d.dist.name, tuple(float(x) for x in d.args))
                    if key not in kinds:
                        kinds.append(key)
                    rew_cls[i, a, k] = kinds.index(key)
        ss = mdp._starting_node_sampler
        tb = cls.from_successors(S, A, succ_idx, succ_prob, succ_len, rew_cls, kinds,
                                 [n2i[n] for n in ss.next_nodes], list(ss.probs),
                                 H=int(mdp.H) if mdp.is_episodic() else 0, rewards_range=tuple(mdp.rewards_range),
                                 T=np.asarray(mdp.T, np.float32), nq=nq)
        tb.R = np.asarray(mdp.R, np.float32)
        tb.node_to_index = dict(n2i)
        tb.index_to_node = {i: n for n, i in n2i.items()}
        return tb

    # ------------------------------------------------------------------------------------------ derived
    @property
    def ld(self):
        """dense row stride in elements.  Rows of up to 1024 states are padded to whole 128-entry chunks (one 128-bit
        load per lane per chunk, no tail predicates in the step kernel); longer rows to a multiple of 32."""
        if self.S <= 1024:
            return (self.S + 127) // 128 * 128
        return (self.S + 31) // 32 * 32

    @property
    def n_start(self):
        return len(self.start_idx)

    @property
    def starting_state_distribution(self):
        """f64 [S], the reference's `starting_state_distribution` (mdp/base.py:494-503)"""
        d = np.zeros(self.S, np.float64)
        np.add.at(d, np.asarray(self.start_idx, np.int64), np.diff(self.start_cum, prepend=0.0))
        return d

    def expected_rewards(self):
        """R[s,a] = sum_s' p * E[r]  (mdp/utils/mdp_creation.py:71-81) from the tables, for cross-checks."""
        import scipy.stats

        means = []
        for k, a in self.rew_kinds:
            means.append(float(a[0]) if k == "deterministic" else float(getattr(scipy.stats, k)(*a).mean()))
        means = np.asarray(means)
        valid = np.arange(self.succ_idx.shape[-1])[None, None, :] < self.succ_len[..., None]
        p = np.diff(np.where(valid, self.succ_cum, 0.0), axis=-1, prepend=0.0)
        p = np.where(valid, p, 0.0)
        return (p * means[self.rew_cls_succ]).sum(-1)
