"""UCRL2Continuous device loops: reward rate over time and throughput (python scripts/ucrl2_probe.py)."""
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

for inst, N, T, kw in (("riverswimcontinuous_ergo0", 256, 30000, dict(alpha_r=0.1, alpha_p=0.05)),
                       ("riverswimcontinuous_ergo0", 256, 30000, dict(alpha_r=1.0, alpha_p=1.0)),
                       ("deepsea10", 256, 30000, dict(alpha_r=0.1, alpha_p=0.05)),
                       ("frozenlakecontinuous_ergo0", 1024, 20000, dict(alpha_r=0.1, alpha_p=0.05, bound_type_p="bernstein")),
                       ("c2_deepsea30_prand", 64, 20000, dict(alpha_r=0.1, alpha_p=0.05))):
    g = load_instance(inst)
    tb = MDPTables.from_golden(g)
    Tm, Rm = np.asarray(g["T"], np.float32), np.asarray(g["R"], np.float32)
    Q, _ = dp.discounted_value_iteration(Tm, Rm)
    opt = markov_chain.get_average_reward(Tm, Rm, dp.get_policy_from_q_values(Q, True))
    rnd = markov_chain.get_average_reward(Tm, Rm, np.full(Rm.shape, 1.0 / tb.A, np.float32))
    ag = al.UCRL2Continuous(0, tb, T + 1, n_loops=N, **kw)
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
    print(f"{inst} S={tb.S} A={tb.A} N={N} T={T} {kw}: optimal {opt:.4f} random {rnd:.4f} reward rate per fifth "
          f"{[round(r, 4) for r in rates]}; {dt:.2f}s, {N * T / dt / 1e6:.2f} M agent-steps/s, {rounds} rounds, "
          f"{ag.evi_iterations} EVI iterations, episodes/loop {float(ag.episode.double().mean()):.0f}", flush=True)
