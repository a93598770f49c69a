"""Throughput of the batched Q-learning loops (agent-steps/s) vs number of loops; CPU oracle port beside it."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
from conftest import load_instance
from colosseum_b200.tables import MDPTables
import colosseum_b200.agent_loop as al
from oracle import oracle as orc
from make_qlearning_golden import host_tables

for inst, mk in (("c2_deepsea30_prand", lambda tb, n: al.QLearningContinuous(0, tb, 10**6, n_loops=n)),
                 ("taxi_epi", lambda tb, n: al.QLearningEpisodic(0, tb, 10**6, p=0.05, c_1=0.5, c_2=0.5, UCB_type="bernstein", n_loops=n)),
                 ("c1_riverswim_epi", lambda tb, n: al.QLearningEpisodic(0, tb, 10**6, p=0.05, c_1=0.5, n_loops=n))):
    tb = MDPTables.from_golden(load_instance(inst))
    for n in (1, 1024, 16384, 65536, 262144):
        per_loop = (tb.H or 1) * tb.S * tb.A * 4 * (5 if tb.H else 3)
        if per_loop * n > 60e9:
            continue
        ag = mk(tb, n)
        K = 500
        ag.steps(50)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ag.steps(K); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"{inst:20s} S={tb.S} A={tb.A} H={tb.H} loops={n:7d}: {ms / K * 1e3:8.2f} us/step  {n * K / ms / 1e6:9.3f} G agent-steps/s "
              f"(tables {per_loop * n / 2**20:.0f} MiB)", flush=True)
        del ag
        torch.cuda.empty_cache()
    kw = dict(optimization_horizon=10**6) if tb.H == 0 else dict(optimization_horizon=10**6, p=0.05, c_1=0.5, c_2=0.5, UCB_type="bernstein" if inst == "taxi_epi" else "hoeffding")
    L = orc.QLearningLoops(host_tables(tb), 4096, seed=0, **kw)
    L.steps(20)
    t0 = time.perf_counter(); L.steps(500); dt = time.perf_counter() - t0
    print(f"{inst:20s} CPU port, 4096 loops, {os.cpu_count()} threads: {4096 * 500 / dt / 1e6:.2f} M agent-steps/s", flush=True)
