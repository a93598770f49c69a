"""Generate the committed golden vectors under tests/golden/ by running the UNMODIFIED Python reference.

Run in the build container only (needs /root/reference; the GPU box has no reference):

    python tests/golden/make_golden.py

What it records (every array is an output of, or an input to, reference code -- nothing here comes from
the product or the oracle):

  inst_<name>.npz   one file per MDP instance:
      T, R                      colosseum/mdp/base.py:943-963 (-> mdp/utils/mdp_creation.py:41-95)
      start_idx, start_prob     colosseum/mdp/base.py:494-503
      succ_idx/succ_prob/succ_len   the NextStateSampler successor lists in construction order
                                (mdp/utils/mdp_creation.py:276-310)
      rew_cls, rew_kinds        per-successor reward distribution id + (scipy name, args) table
                                (mdp/base.py:1170-1185)
      vi_Q, vi_V                reference VI at its defaults (episodic: finite_horizon.py:11-26;
                                continuous: infinite_horizon.py:14-44 eps=1e-3)
      vi_tight_Q, vi_tight_V    continuous only: reference numba VI with eps=1e-6 (near fixed point)
      pe_Q, pe_V                reference policy evaluation of the uniform random policy
      diameter, value_norm, gaps        mdp.diameter / mdp.value_norm / mdp.sum_reciprocals_suboptimality_gaps
      diameter_tight            continuous, S<=110: the reference's per-target VI (diameter.py:76-95) with eps=2e-5
      cached_diameter, cached_value_norm    the reference's own cached_hardness_measures/*.txt (NaN if absent)
      T_epi (episodic)          mdp/utils/mdp_creation.py:98-128
      T_cf, R_cf, vi_cf_V (episodic)    continuous form + its VI (mdp/base_finite.py:167-178)
      reach_h, reach_s (episodic)       mdp/base_finite.py:138-150
      traj_*                    a reset + n-step reference trajectory with the uniforms its samplers consumed
  sampler_kat.npz   (probs, uniforms, chosen position) for raw NextStateSampler instances
                    (mdp/utils/custom_samplers.py:49-72 == CPython random.choices)
  dp_synth.npz      reference numba DP on small synthetic dense MDPs (config C4's generator at S=24)
  doc_goldens.json  numbers printed in the reference's executed notebooks
"""
import json
import os
import random
import re
import sys
from bisect import bisect as _bisect

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.reference_import import REFERENCE_ROOT, import_reference  # noqa: E402

colosseum = import_reference()
import colosseum.mdp  # noqa: E402,F401  (must precede colosseum.hardness: the reference has an L1<->L2 import cycle)

from colosseum.dynamic_programming import (  # noqa: E402
    discounted_value_iteration,
    episodic_policy_evaluation,
    episodic_value_iteration,
)
from colosseum.dynamic_programming.infinite_horizon import (  # noqa: E402
    _discounted_policy_evaluation,
    _discounted_value_iteration,
)
from colosseum.hardness.measures import (  # noqa: E402
    calculate_norm_discounted,
    get_diameter,
    get_sum_reciprocals_suboptimality_gaps,
)
from colosseum.mdp.utils import custom_samplers  # noqa: E402

# ---- record the seed every NextStateSampler is built with (runtime patch; reference files untouched) ----
_orig_init = custom_samplers.NextStateSampler.__init__


def _init_recording_seed(self, next_nodes, seed=None, probs=None):
    self._golden_seed = seed
    _orig_init(self, next_nodes, seed=seed, probs=probs)


custom_samplers.NextStateSampler.__init__ = _init_recording_seed

CACHE = os.path.join(REFERENCE_ROOT, "colosseum", "benchmark", "cached_hardness_measures")


def gin_param_sets(gin_path, cls_name):
    """20-line reader for the reference's `prms_i/Class.key=value` gin files."""
    sets = {}
    with open(gin_path) as f:
        for line in f:
            m = re.match(r"\s*prms_(\d+)/" + cls_name + r"\.(\w+)\s*=\s*(.+?)\s*$", line)
            if m:
                sets.setdefault(int(m.group(1)), {})[m.group(2)] = eval(m.group(3))  # trusted local file
    return [sets[k] for k in sorted(sets)]


def cached_measure(mdp, measure):
    p = os.path.join(CACHE, type(mdp).__name__, f"{measure}_{mdp.hash}.txt")
    if os.path.isfile(p):
        txt = open(p).read().strip()
        if txt:
            return float(txt)
    return float("nan")


def sampler_uniform_stream(sampler, n):
    """The first n uniforms the sampler's private random.Random(seed) produced/will produce."""
    twin = random.Random(sampler._golden_seed)
    return [twin.random() for _ in range(n)]


def inv_cdf_position(probs, u):
    """CPython random.choices restated: bisect_right(accumulate(probs), u*total, 0, n-1)."""
    cum = []
    acc = None
    for p in probs:
        acc = p if acc is None else acc + p
        cum.append(acc)
    total = cum[-1] + 0.0
    return _bisect(cum, u * total, 0, len(probs) - 1)


def reference_diameter_tight(T, eps=2e-5):
    """the reference's per-target VI (hardness/measures/diameter.py:76-95) run with a small epsilon instead of its
    default 1e-3, so that the recorded value sits at the fixed point instead of at the early stop."""
    from colosseum.dynamic_programming.utils import DynamicProgrammingMaxIterationExceeded

    S = T.shape[0]
    best = 0.0
    for es in range(S):
        T_es = T.copy()
        T_es[es] = 0
        T_es[es, :, es] = 1
        R_es = np.zeros(T.shape[:2], np.float32) - 1.0
        R_es[es] = 0
        e = eps
        while True:
            try:
                _, V = _discounted_value_iteration(T_es, R_es, 1.0, e)
                break
            except DynamicProgrammingMaxIterationExceeded:
                e *= 4
        best = max(best, float(-V.min()))
    return best


def dist_key(d):
    name = d.dist.name
    args = tuple(float(a) for a in d.args)
    return (name, args)


def dump_instance(name, mdp, traj_steps=400, do_diameter=True, traj_seed=7):
    S, A = mdp.n_states, mdp.n_actions
    episodic = bool(mdp.is_episodic())
    T, R = mdp.T, mdp.R
    out = dict(
        T=T,
        R=R,
        n_states=S,
        n_actions=A,
        episodic=episodic,
        H=int(mdp.H) if episodic else 0,
        rewards_range=np.asarray(mdp.rewards_range, np.float64),
        cls_name=type(mdp).__name__,
        mdp_hash=mdp.hash,
    )
    n2i = mdp.node_to_index
    nodes = list(mdp.G.nodes)
    assert [n2i[n] for n in nodes] == list(range(S))

    # ---- starting distribution and successor lists, in the samplers' own order ----
    ss = mdp._starting_node_sampler
    out["start_idx"] = np.asarray([n2i[n] for n in ss.next_nodes], np.int32)
    out["start_prob"] = np.asarray(ss.probs, np.float64)
    K = max(len(mdp.get_info_class(n).transition_distributions[a].next_nodes) for n in nodes for a in range(A))
    succ_idx = -np.ones((S, A, K), np.int32)
    succ_prob = np.zeros((S, A, K), np.float64)
    succ_len = np.zeros((S, A), np.int32)
    rew_cls = -np.ones((S, A, K), np.int32)
    kinds = []
    for i, n in enumerate(nodes):
        for a in range(A):
            td = mdp.get_info_class(n).transition_distributions[a]
            succ_len[i, a] = len(td.next_nodes)
            for k, (nn, p) in enumerate(zip(td.next_nodes, td.probs)):
                succ_idx[i, a, k] = n2i[nn]
                succ_prob[i, a, k] = p
                key = dist_key(mdp.get_reward_distribution(n, a, nn))
                if key not in kinds:
                    kinds.append(key)
                rew_cls[i, a, k] = kinds.index(key)
    out.update(succ_idx=succ_idx, succ_prob=succ_prob, succ_len=succ_len, rew_cls=rew_cls)
    out["rew_kinds"] = json.dumps(kinds)

    # ---- reference DP ----
    pi_rand = np.ones((S, A), np.float32) / A
    if episodic:
        H = mdp.H
        Q, V = episodic_value_iteration(H, T, R)
        out.update(vi_Q=Q, vi_V=V)
        pol = np.ones((H, S, A), np.float32) / A
        Q, V = episodic_policy_evaluation(H, T, R, pol)
        out.update(pe_Q=Q, pe_V=V)
        T_epi, R_epi = mdp.episodic_transition_matrix_and_rewards
        out["T_epi"] = T_epi
        T_cf, R_cf = mdp.T_cf, mdp.R_cf
        out.update(T_cf=T_cf, R_cf=R_cf, vi_cf_V=mdp.optimal_value_continuous_form[1])
        rs = mdp.reachable_states
        out.update(reach_h=np.asarray([h for h, _ in rs], np.int32), reach_s=np.asarray([s for _, s in rs], np.int32))
    else:
        Q, V = discounted_value_iteration(T, R)
        out.update(vi_Q=Q, vi_V=V)
        Q, V = _discounted_value_iteration(T, R, 0.99, 1e-6)
        out.update(vi_tight_Q=Q, vi_tight_V=V)
        # dense numba branch directly (the dispatcher's COO branch needs pydata/sparse, absent here)
        Q, V = _discounted_policy_evaluation(T, R, pi_rand, 0.99, 1e-7)
        out.update(pe_Q=Q, pe_V=V)

    # ---- reference hardness measures ----
    out["value_norm"] = float(mdp.value_norm)
    out["all_deterministic"] = bool(mdp._are_all_transition_deterministic and mdp._are_all_rewards_deterministic)
    out["gaps"] = float(mdp.sum_reciprocals_suboptimality_gaps)
    out["diameter"] = float(mdp.diameter) if do_diameter else float("nan")
    out["cached_diameter"] = cached_measure(mdp, "diameter")
    out["cached_value_norm"] = cached_measure(mdp, "value_norm")
    if not episodic:
        out["norm_of_vi_V"] = float(calculate_norm_discounted(T, out["vi_V"]))
        out["diameter_tight"] = reference_diameter_tight(T) if S <= 110 else float("nan")

    # ---- a reference trajectory, with the uniforms its samplers consumed ----
    rng = np.random.RandomState(traj_seed)
    counters = {}
    streams = {}

    def next_uniform(sampler):
        if sampler.is_deterministic:
            return float("nan")
        sid = id(sampler)
        if sid not in streams:
            # the sampler drew 5000 at construction and draws 5000 more per refill, all from one stream
            streams[sid] = sampler_uniform_stream(sampler, 5000 + 5000 * (2 + traj_steps // 5000))
            counters[sid] = 0
        u = streams[sid][counters[sid]]
        counters[sid] += 1
        return u

    mdp.reset_visitation_counts()
    t_action, t_u, t_type, t_obs, t_rew, t_disc = [], [], [], [], [], []
    u0 = next_uniform(ss)
    ts = mdp.reset()
    t_action.append(-1); t_u.append(u0); t_type.append(int(ts.step_type)); t_obs.append(int(ts.observation))
    t_rew.append(float("nan")); t_disc.append(float("nan"))
    for _ in range(traj_steps):
        a = int(rng.randint(A))
        if mdp.necessary_reset:  # auto_reset path: the step call is a reset and ignores the action
            u = next_uniform(ss)
        else:
            u = next_uniform(mdp.get_info_class(mdp.cur_node).transition_distributions[a])
        ts = mdp.step(a, auto_reset=True)
        t_action.append(a); t_u.append(u); t_type.append(int(ts.step_type)); t_obs.append(int(ts.observation))
        t_rew.append(float("nan") if ts.reward is None else float(ts.reward))
        t_disc.append(float("nan") if ts.discount is None else float(ts.discount))
    vis_s = np.asarray([mdp.get_info_class(n).state_visitation_count for n in nodes], np.int64)
    vis_sa = np.asarray(
        [[mdp.get_info_class(n).actions_visitation_count[a] for a in range(A)] for n in nodes], np.int64
    )
    out.update(
        traj_action=np.asarray(t_action, np.int32),
        traj_u=np.asarray(t_u, np.float64),
        traj_step_type=np.asarray(t_type, np.int8),
        traj_obs=np.asarray(t_obs, np.int32),
        traj_reward=np.asarray(t_rew, np.float64),
        traj_discount=np.asarray(t_disc, np.float64),
        traj_visits_s=vis_s,
        traj_visits_sa=vis_sa,
    )
    np.savez_compressed(os.path.join(HERE, f"inst_{name}.npz"), **out)
    print(
        f"{name}: S={S} A={A} H={out['H']} K={K} diam={out['diameter']:.6f} (cached {out['cached_diameter']:.6f}) "
        f"vnorm={out['value_norm']:.6f} (cached {out['cached_value_norm']:.6f}) gaps={out['gaps']:.4f}"
    )


def sampler_kat():
    rs = np.random.RandomState(0)
    recs = dict()
    for i, n in enumerate([2, 3, 5, 10, 17]):
        p = rs.dirichlet(np.ones(n) * 0.5)
        if i == 2:
            p[1] = 0.0  # a zero-probability successor
            p = p / p.sum()
        probs = [float(x) for x in p]
        s = custom_samplers.NextStateSampler(next_nodes=list(range(100, 100 + n)), seed=1000 + i, probs=probs)
        us = sampler_uniform_stream(s, 5000)
        chosen = np.asarray([c - 100 for c in s.cached_states], np.int32)
        mine = np.asarray([inv_cdf_position(probs, u) for u in us], np.int32)
        assert (mine == chosen).all(), "inverse-CDF restatement disagrees with CPython random.choices"
        recs[f"probs_{i}"] = np.asarray(probs, np.float64)
        recs[f"u_{i}"] = np.asarray(us, np.float64)
        recs[f"chosen_{i}"] = chosen
    np.savez_compressed(os.path.join(HERE, "sampler_kat.npz"), **recs)
    print("sampler_kat: CPython random.choices == bisect restatement on 5 x 5000 draws")


def dp_synth():
    """Reference numba DP on config-C4-style synthetic MDPs (Dirichlet(0.05) rows), small S."""
    recs = {}
    for b, (S, A) in enumerate([(24, 4), (33, 3), (64, 2)]):
        rs = np.random.RandomState(b)
        T = rs.dirichlet(np.ones(S) * 0.05, size=(S, A)).astype(np.float32)
        T = (T / T.sum(-1, keepdims=True, dtype=np.float32)).astype(np.float32)
        R = rs.uniform(0, 1, size=(S, A)).astype(np.float32)
        pi = rs.dirichlet(np.ones(A), size=S).astype(np.float32)
        Q, V = _discounted_value_iteration(T, R, 0.99, 1e-3)
        Qt, Vt = _discounted_value_iteration(T, R, 0.99, 1e-6)
        Qp, Vp = _discounted_policy_evaluation(T, R, pi, 0.99, 1e-7)
        H = 7
        Qe, Ve = episodic_value_iteration(H, T, R)
        pol = rs.dirichlet(np.ones(A), size=(H, S)).astype(np.float32)
        Qpe, Vpe = episodic_policy_evaluation(H, T, R, pol)
        d = get_diameter(T, False)
        recs[f"diam_tight_{b}"] = reference_diameter_tight(T)
        vn = calculate_norm_discounted(T, V)
        gaps = get_sum_reciprocals_suboptimality_gaps(Q, V)
        recs.update({
            f"T_{b}": T, f"R_{b}": R, f"pi_{b}": pi, f"Q_{b}": Q, f"V_{b}": V, f"Qt_{b}": Qt, f"Vt_{b}": Vt,
            f"Qp_{b}": Qp, f"Vp_{b}": Vp, f"H_{b}": H, f"Qe_{b}": Qe, f"Ve_{b}": Ve, f"pol_{b}": pol,
            f"Qpe_{b}": Qpe, f"Vpe_{b}": Vpe, f"diam_{b}": float(d), f"vnorm_{b}": float(vn), f"gaps_{b}": float(gaps),
        })
        # overflow contract: returns None (colosseum/dynamic_programming/infinite_horizon.py:136-138)
        assert _discounted_value_iteration(T, R, 0.99, 1e-3, 5.0) is None
        assert episodic_value_iteration(H, T, R, 1.5) is None
        print(f"dp_synth[{b}]: S={S} A={A} diam={d:.5f} vnorm={vn:.6f} gaps={gaps:.4f}")
    np.savez_compressed(os.path.join(HERE, "dp_synth.npz"), **recs)


def main(only=None):
    global dump_instance
    if only:
        _dump = dump_instance

        def dump_instance(name, *a, **k):  # noqa: F811
            if name in only:
                _dump(name, *a, **k)

    from colosseum.mdp.deep_sea import DeepSeaContinuous, DeepSeaEpisodic
    from colosseum.mdp.frozen_lake import FrozenLakeContinuous, FrozenLakeEpisodic
    from colosseum.mdp.minigrid_empty import MiniGridEmptyContinuous, MiniGridEmptyEpisodic
    from colosseum.mdp.minigrid_rooms import MiniGridRoomsContinuous
    from colosseum.mdp.river_swim import RiverSwimContinuous, RiverSwimEpisodic
    from colosseum.mdp.simple_grid import SimpleGridContinuous, SimpleGridEpisodic
    from colosseum.mdp.taxi import TaxiContinuous, TaxiEpisodic

    bench = os.path.join(REFERENCE_ROOT, "colosseum", "benchmark")
    if not only:
        sampler_kat()
        dp_synth()

    # C1: the quick-test RiverSwimEpisodic (BASELINE.json configs[0]); V[0] is the SURVEY parity anchor
    (p,) = gin_param_sets(
        os.path.join(bench, "benchmark_episodic_quick_test", "mdp_configs", "RiverSwimEpisodic.gin"), "RiverSwimEpisodic"
    )
    dump_instance("c1_riverswim_epi", RiverSwimEpisodic(seed=0, **p))

    # documentation notebook instance (hardness-analysis.ipynb)
    dump_instance("doc_simplegrid4", SimpleGridContinuous(seed=0, size=4, p_rand=0.01, n_starting_states=3))

    # continuous-class instances whose (class, params, seed) have files in the reference's hardness cache
    dump_instance("deepsea20_prand", DeepSeaContinuous(seed=0, size=20, p_rand=0.1))
    dump_instance("deepsea10", DeepSeaContinuous(seed=0, size=10, p_rand=None))
    # C2 (BASELINE.json configs[1]): DeepSeaContinuous size 30, both variants of SURVEY.md section 8d; the
    # reference's diameter takes minutes at S=465 and is not recorded
    dump_instance("c2_deepsea30", DeepSeaContinuous(seed=0, size=30, p_rand=None), do_diameter=False)
    dump_instance("c2_deepsea30_prand", DeepSeaContinuous(seed=0, size=30, p_rand=0.1), do_diameter=False)
    for cls, gin_dir, picks in [
        (FrozenLakeContinuous, "benchmark_continuous_ergodic", [0]),
        (FrozenLakeContinuous, "benchmark_continuous_communicating", [0]),
        (TaxiContinuous, "benchmark_continuous_ergodic", [0]),
        (RiverSwimContinuous, "benchmark_continuous_ergodic", [0, 1]),
        (SimpleGridContinuous, "benchmark_continuous_ergodic", [1]),
        (MiniGridEmptyContinuous, "benchmark_continuous_ergodic", [0, 1]),
        (MiniGridRoomsContinuous, "benchmark_continuous_ergodic", [0]),
    ]:
        sets = gin_param_sets(os.path.join(bench, gin_dir, "mdp_configs", cls.__name__ + ".gin"), cls.__name__)
        for i in picks:
            if i >= len(sets):
                continue
            mdp = cls(seed=0, **sets[i])
            if mdp.n_states > 420:
                print(f"skip {cls.__name__}[{gin_dir},{i}]: S={mdp.n_states}")
                continue
            tag = gin_dir.split("_")[-1][:4]
            dump_instance(f"{cls.__name__.lower()}_{tag}{i}", mdp)

    # episodic instances (the episodic cache files are stale -- SURVEY.md section 4 -- so reference-computed only)
    dump_instance("deepsea8_epi", DeepSeaEpisodic(seed=0, size=8, p_rand=0.2))
    dump_instance("frozenlake4_epi", FrozenLakeEpisodic(seed=1, size=4, p_frozen=0.8))
    dump_instance("taxi_epi", TaxiEpisodic(seed=0, size=4, length=1, width=1, space=1, n_locations=3, p_rand=0.1),
                  do_diameter=True)
    dump_instance("simplegrid5_epi", SimpleGridEpisodic(seed=2, size=5, p_lazy=0.1, p_rand=0.1))
    dump_instance("minigridempty5_epi", MiniGridEmptyEpisodic(seed=0, size=5, p_rand=0.1))

    with open(os.path.join(HERE, "doc_goldens.json"), "w") as f:
        json.dump(
            {
                "source": "docs/_sources/mds/hardness-analysis.ipynb:83-193 (SimpleGridContinuous seed=0 size=4 p_rand=0.01 n_starting_states=3)",
                "diameter": 6.0545096,
                "value_norm": 0.49540126,
                "suboptimal_gaps": 361.29538,
                "c1_V0": [0.45454547, 0.36414355, 0.2737823, 0.3346822, 0.4166667],
            },
            f,
            indent=1,
        )


if __name__ == "__main__":
    main(only=set(sys.argv[1:]))  # optional: names of the instances to (re)generate
