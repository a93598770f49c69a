"""PSRLContinuous device loops: reward rate over time and throughput (python scripts/psrlc_probe.py)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_instance  # noqa: E402
import colosseum_b200.agent_loop as al  # noqa: E402
from colosseum_b200 import dynamic_programming as dp, markov_chain  # noqa: E402
from colosseum_b200.tables import MDPTables  # noqa: E402

for inst, N, T, kw in (("riverswimcontinuous_ergo0", 256, 20000, dict(psi_weight=0.015, eta_weight=1e-9)),
                       ("riverswimcontinuous_ergo0", 256, 20000, dict(no_optimistic_sampling=True)),
                       ("deepsea10", 256, 20000, dict(psi_weight=0.01, eta_weight=1e-9)),
                       ("frozenlakecontinuous_ergo0", 512, 20000, dict(psi_weight=0.02, eta_weight=1e-8)),
                       ("c2_deepsea30_prand", 32, 5000, dict(psi_weight=0.0005, eta_weight=1e-10))):
    g = load_instance(inst)
    tb = MDPTables.from_golden(g)
    Tm, Rm = np.asarray(g["T"], np.float32), np.asarray(g["R"], np.float32)
    Q, _ = dp.discounted_value_iteration(Tm, Rm)
    opt = markov_chain.get_average_reward(Tm, Rm, dp.get_policy_from_q_values(Q, True))
    for so in ("gauss_seidel", "jacobi"):
        ag = al.PSRLContinuous(0, tb, T + 1, n_loops=N, sweep_order=so, **kw)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rates, prev, rounds = [], 0.0, 0
        for _ in range(5):
            ag.steps(T // 5)
            rounds += ag.rounds
            cum = float(ag.cumulative_reward.mean())
            rates.append((cum - prev) / (T // 5))
            prev = cum
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"{inst} S={tb.S} A={tb.A} psi={ag._psi} eta={ag._eta:.3g} N={N} T={T} {so}: optimal {opt:.4f} reward rate per "
              f"fifth {[round(r, 4) for r in rates]}; {dt:.2f}s, {N * T / dt / 1e6:.2f} M agent-steps/s, {rounds} rounds, "
              f"{ag.vi_sweeps} VI sweeps, re-plannings/loop {float(ag.episode.double().mean()):.0f}", flush=True)
