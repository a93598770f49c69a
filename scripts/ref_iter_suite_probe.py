import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import colosseum_b200.dynamic_programming as dp
import colosseum_b200.hardness as hd
from colosseum_b200.suite import load_suite
rel = lambda a, b: abs(a - b) / max(abs(b), 1e-12)
worst = {"gaps": 0, "vn": 0, "diam": 0}
for inst in load_suite("tests/golden/c3_suite.npz"):
    if inst.episodic: continue
    T = torch.from_numpy(inst.tables.T).cuda(); R = torch.from_numpy(inst.R).cuda()
    t0 = time.perf_counter()
    Q, V = dp.discounted_value_iteration(T, R, sweep_order="gauss_seidel")   # reference defaults
    gaps = hd.get_sum_reciprocals_suboptimality_gaps(Q, V)
    det = bool((inst.tables.succ_len == 1).all()) and all(k == "deterministic" for k, _ in inst.tables.rew_kinds)
    vn = 0.0 if det else hd.calculate_norm_discounted(T, V, precision="f32")
    d = hd.get_diameter(T, False, reference_iterates=True)
    dt = time.perf_counter() - t0
    r = inst.ref
    eg, ev = rel(gaps, r["gaps"]), abs(vn - r["value_norm"]) / max(r["value_norm"], 0.05)
    ed = abs(d - r["diameter"]) if np.isfinite(d) and np.isfinite(r["diameter"]) else float("nan")
    worst["gaps"] = max(worst["gaps"], eg); worst["vn"] = max(worst["vn"], ev)
    if np.isfinite(ed): worst["diam"] = max(worst["diam"], ed)
    print(f"{inst.name:34s} S={inst.S:4d} {dt*1e3:7.1f} ms gaps rel {eg:.1e} vn rel {ev:.1e} diam abs {ed:.1e}")
print("worst", worst)
