"""Generate tests/golden/c3_suite.npz: the reference's benchmark MDP instances (BASELINE.json configs[2], "C3") in
SPARSE form, by running the UNMODIFIED Python reference in the build container (needs /root/reference).

    python tests/golden/make_c3_suite.py [max_seconds_per_reference_diameter]

For every `prms_i` parameter set of every family gin file under
colosseum/benchmark/benchmark_{continuous,episodic}_{ergodic,communicating}/mdp_configs/ (seed 0) it records what the
host-side table extractor sees -- the NextStateSampler successor lists in construction order
(mdp/utils/mdp_creation.py:276-310), the reward distribution per successor (mdp/base.py:1170-1185), the start
distribution (mdp/base.py:494-503), H and rewards_range -- plus the reference's own answers: R (mdp/base.py:943-963),
value_norm / gaps / diameter (mdp/base.py:996-1114; the diameter only when it is cached by the reference or cheap),
and a CRC of mdp.T so the loader can prove that the dense T it rebuilds from the successor lists is bit-identical.
Dense T tensors (up to 11 MB each) are NOT stored: the file stays small enough to commit.
"""
import json
import os
import sys
import time
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from make_golden import CACHE, cached_measure, dist_key, gin_param_sets  # noqa: E402,F401  (imports the reference)
from oracle.reference_import import REFERENCE_ROOT  # noqa: E402

import colosseum.mdp as cmdp  # noqa: E402

BENCH = os.path.join(REFERENCE_ROOT, "colosseum", "benchmark")
FAMILIES = ["DeepSea", "FrozenLake", "MiniGridEmpty", "MiniGridRooms", "RiverSwim", "SimpleGrid", "Taxi"]


def family_class(fam, kind):
    import importlib

    mod = {"DeepSea": "deep_sea", "FrozenLake": "frozen_lake", "MiniGridEmpty": "minigrid_empty",
           "MiniGridRooms": "minigrid_rooms", "RiverSwim": "river_swim", "SimpleGrid": "simple_grid", "Taxi": "taxi"}[fam]
    m = importlib.import_module(f"colosseum.mdp.{mod}")
    return getattr(m, fam + kind)


def sparse_tables(mdp):
    S, A = mdp.n_states, mdp.n_actions
    n2i = mdp.node_to_index
    nodes = list(mdp.G.nodes)
    tds = [[mdp.get_info_class(n).transition_distributions[a] for a in range(A)] for n in nodes]
    K = max(len(td.next_nodes) for row in tds for td in row)
    succ_idx = np.zeros((S, A, K), np.int32)
    succ_prob = np.zeros((S, A, K), np.float64)
    succ_len = np.zeros((S, A), np.int8)
    rew_cls = np.zeros((S, A, K), np.int8)
    kinds = []
    for i, n in enumerate(nodes):
        for a in range(A):
            td = tds[i][a]
            succ_len[i, a] = len(td.next_nodes)
            for k, (nn, p) in enumerate(zip(td.next_nodes, td.probs)):
                succ_idx[i, a, k] = n2i[nn]
                succ_prob[i, a, k] = p
                key = dist_key(mdp.get_reward_distribution(n, a, nn))
                if key not in kinds:
                    kinds.append(key)
                rew_cls[i, a, k] = kinds.index(key)
    ss = mdp._starting_node_sampler
    return dict(succ_idx=succ_idx.astype(np.int16 if S < 32768 else np.int32), succ_prob=succ_prob, succ_len=succ_len,
                rew_cls=rew_cls, rew_kinds=json.dumps(kinds),
                start_idx=np.asarray([n2i[n] for n in ss.next_nodes], np.int32),
                start_prob=np.asarray(ss.probs, np.float64))


def main(max_diam_seconds=20.0):
    out = {}
    names = []
    t_all = time.time()
    for kind, benches in (("Continuous", ["benchmark_continuous_ergodic", "benchmark_continuous_communicating"]),
                          ("Episodic", ["benchmark_episodic_ergodic", "benchmark_episodic_communicating"])):
        for bench in benches:
            for fam in FAMILIES:
                cls = family_class(fam, kind)
                sets = gin_param_sets(os.path.join(BENCH, bench, "mdp_configs", cls.__name__ + ".gin"), cls.__name__)
                for i, prm in enumerate(sets):
                    t0 = time.time()
                    mdp = cls(seed=0, **prm)
                    S, A = mdp.n_states, mdp.n_actions
                    episodic = bool(mdp.is_episodic())
                    name = f"{cls.__name__}.{bench.split('_')[-1][:4]}{i}"
                    key = f"i{len(names)}_"
                    tb = sparse_tables(mdp)
                    T = np.ascontiguousarray(mdp.T, np.float32)
                    rec = dict(tb, S=S, A=A, H=int(mdp.H) if episodic else 0,
                               rewards_range=np.asarray(mdp.rewards_range, np.float64), R=np.asarray(mdp.R, np.float32),
                               T_crc=np.uint32(zlib.crc32(T.tobytes())), params=json.dumps(prm, default=str),
                               mdp_hash=mdp.hash)
                    # reference hardness answers (value_norm / gaps need one VI; the episodic value norm needs the
                    # O(nodes^2) continuous-form builder, skipped when the (h,s) graph is large)
                    try:
                        rec["gaps"] = float(mdp.sum_reciprocals_suboptimality_gaps)
                    except Exception as e:  # pragma: no cover
                        print("   gaps failed:", e)
                        rec["gaps"] = float("nan")
                    small_cf = (not episodic) or S * mdp.H <= 6000
                    if episodic:
                        # node order of the reference's episodic graph (base_finite.py:138-150): the order of the rows
                        # of T_cf, which -- through the start-column quirk of mdp_creation.py:168 -- the reference's
                        # episodic value norm depends on
                        rs = mdp.reachable_states
                        rec["reach_h"] = np.asarray([h for h, _ in rs], np.int16)
                        rec["reach_s"] = np.asarray([s_ for _, s_ in rs], np.int16)
                    rec["value_norm"] = float(mdp.value_norm) if small_cf else float("nan")
                    rec["cached_diameter"] = cached_measure(mdp, "diameter") if not episodic else float("nan")
                    rec["cached_value_norm"] = cached_measure(mdp, "value_norm") if not episodic else float("nan")
                    est = (S / 100.0) ** 3 * (3.0 if not episodic else 0.5 * mdp.H)  # crude seconds estimate
                    if est <= max_diam_seconds:
                        td = time.time()
                        rec["diameter"] = float(mdp.diameter)
                        rec["diameter_seconds"] = time.time() - td
                    else:
                        rec["diameter"] = float("nan")
                        rec["diameter_seconds"] = float("nan")
                    for k, v in rec.items():
                        out[key + k] = v
                    names.append(name)
                    print(f"[{len(names):3d}] {name:40s} S={S:4d} A={A} H={rec['H']:3d} K={tb['succ_idx'].shape[-1]:2d} "
                          f"vn={rec['value_norm']:.5f} gaps={rec['gaps']:.3f} diam={rec['diameter']:.4f} "
                          f"(cached {rec['cached_diameter']:.4f}) {time.time() - t0:.1f}s", flush=True)
    out["names"] = json.dumps(names)
    np.savez_compressed(os.path.join(HERE, "c3_suite.npz"), **out)
    print(f"{len(names)} instances, {os.path.getsize(os.path.join(HERE, 'c3_suite.npz')) / 1e6:.2f} MB, "
          f"{time.time() - t_all:.0f}s")


if __name__ == "__main__":
    main(float(sys.argv[1]) if len(sys.argv) > 1 else 20.0)
