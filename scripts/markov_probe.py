import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import colosseum_b200.markov_chain as mc
from oracle import oracle as orc
g = np.load("tests/golden/avg_reward.npz")
for name in g["names"]:
    T, R = g[f"{name}_T"], g[f"{name}_R"]
    starts = list(zip(g[f"{name}_start_idx"].tolist(), g[f"{name}_start_prob"].tolist()))
    for k in ("opt", "worst", "rand"):
        pi = g[f"{name}_{k}_pi"]
        tps = mc.get_transition_probabilities(T, pi)
        for tol in (1e-8, 1e-10, 1e-12):
            t0 = time.perf_counter()
            try:
                sd = mc.get_stationary_distribution(tps, starts, tol=tol, max_iter=2000000)
                it = mc.get_stationary_distribution.last_iterations
                Po, ro = orc.policy_chain(T, R, pi)
                x0 = np.zeros(len(T)); x0[g[f"{name}_start_idx"]] = g[f"{name}_start_prob"]
                err = np.abs(sd - orc.stationary_distribution_f64(Po, x0)).max()
                print(f"{name:18s} {k:5s} tol={tol:g} iters={it:8d} {1e3*(time.perf_counter()-t0):8.1f} ms err_vs_oracle={err:.2e} ref_err={np.abs(sd-g[f'{name}_{k}_sd']).max():.2e}")
            except Exception as e:
                print(f"{name:18s} {k:5s} tol={tol:g} FAILED {type(e).__name__} after {1e3*(time.perf_counter()-t0):.0f} ms")
