"""Extend the C3 fixture to SURVEY section 8d's full list: the gin parameter sets of ALL six benchmark folders under
colosseum/benchmark/ (the four benchmarks of tests/golden/c3_suite.npz plus the two quick-test ones: 94 parameter
sets) x seeds 0..10, by running the UNMODIFIED Python reference in the build container:

    python tests/golden/make_c3_suite_seeds.py      ->  tests/golden/c3_suite_seeds.npz

c3_suite.npz already holds the 80 (benchmark, parameter set) pairs at seed 0 with the reference's hardness answers;
this file adds the 14 quick-test sets at seed 0 and all 94 sets at seeds 1..10 (954 instances), in SPARSE form as
there (successor lists, reward classes, start distribution, H, R, CRC of mdp.T) with the cheap reference answers:
`gaps` (one VI) and, for continuous classes, whatever the reference's own hardness cache holds for that
(class, parameters, seed).  Diameters / episodic value norms of the reference itself take seconds to minutes per
instance and are recorded for seed 0 only (c3_suite.npz)."""
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

from make_c3_suite import BENCH, FAMILIES, family_class, sparse_tables  # noqa: E402  (imports the reference)
from make_golden import cached_measure, gin_param_sets  # noqa: E402

BENCHES = [("Continuous", "benchmark_continuous_ergodic"), ("Continuous", "benchmark_continuous_communicating"),
           ("Episodic", "benchmark_episodic_ergodic"), ("Episodic", "benchmark_episodic_communicating"),
           ("Continuous", "benchmark_continuous_quick_test"), ("Episodic", "benchmark_episodic_quick_test")]


def main():
    out, names = {}, []
    t_all = time.time()
    for seed in range(0, 11):
        for kind, bench in BENCHES:
            if seed == 0 and "quick_test" not in bench:
                continue  # already in c3_suite.npz
            for fam in FAMILIES:
                cls = family_class(fam, kind)
                sets = gin_param_sets(os.path.join(BENCH, bench, "mdp_configs", cls.__name__ + ".gin"), cls.__name__)
                for i, prm in enumerate(sets):
                    t0 = time.time()
                    mdp = cls(seed=seed, **prm)
                    S, A = mdp.n_states, mdp.n_actions
                    episodic = bool(mdp.is_episodic())
                    tag = "quick" if "quick_test" in bench else bench.split("_")[-1][:4]
                    name = f"{cls.__name__}.{tag}{i}.s{seed}"
                    key = f"i{len(names)}_"
                    tb = sparse_tables(mdp)
                    T = np.ascontiguousarray(mdp.T, np.float32)
                    rec = dict(tb, S=S, A=A, H=int(mdp.H) if episodic else 0,
                               rewards_range=np.asarray(mdp.rewards_range, np.float64), R=np.asarray(mdp.R, np.float32),
                               T_crc=np.uint32(zlib.crc32(T.tobytes())), seed=seed)
                    try:
                        rec["gaps"] = float(mdp.sum_reciprocals_suboptimality_gaps)
                    except Exception as e:  # pragma: no cover
                        print("   gaps failed:", e)
                        rec["gaps"] = float("nan")
                    if episodic:
                        rs = mdp.reachable_states
                        rec["reach_h"] = np.asarray([h for h, _ in rs], np.int16)
                        rec["reach_s"] = np.asarray([s_ for _, s_ in rs], np.int16)
                    rec["value_norm"] = float("nan")
                    rec["diameter"] = float("nan")
                    rec["cached_diameter"] = cached_measure(mdp, "diameter") if not episodic else float("nan")
                    rec["cached_value_norm"] = cached_measure(mdp, "value_norm") if not episodic else float("nan")
                    for k, v in rec.items():
                        out[key + k] = v
                    names.append(name)
                    print(f"[{len(names):4d}] {name:48s} S={S:4d} A={A} H={rec['H']:3d} gaps={rec['gaps']:.3f} "
                          f"cached d={rec['cached_diameter']:.4f} vn={rec['cached_value_norm']:.5f} {time.time() - t0:.1f}s",
                          flush=True)
    out["names"] = json.dumps(names)
    p = os.path.join(HERE, "c3_suite_seeds.npz")
    np.savez_compressed(p, **out)
    print(f"{len(names)} instances, {os.path.getsize(p) / 1e6:.2f} MB, {time.time() - t_all:.0f}s")


if __name__ == "__main__":
    main()
