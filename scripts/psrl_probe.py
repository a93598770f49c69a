"""Throughput of the batched PSRL loops: steps kernel vs per-episode resample + VI."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import load_instance
from colosseum_b200.tables import MDPTables
import colosseum_b200.agent_loop as al

for inst in ("c1_riverswim_epi", "taxi_epi", "deepsea8_epi"):
    tb = MDPTables.from_golden(load_instance(inst))
    for n in (64, 1024, 8192):
        if n * tb.S * tb.A * tb.S * 8 > 40e9:
            continue
        ag = al.PSRLEpisodic(0, tb, 10 ** 6, n_loops=n)
        ag.steps(2 * tb.H)
        torch.cuda.synchronize()
        E = 20
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ag.steps(E * tb.H); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        e0.record()
        for _ in range(E): ag.episode_end_update()
        e1.record(); torch.cuda.synchronize()
        ms_s = e0.elapsed_time(e1)
        print(f"{inst:18s} S={tb.S} A={tb.A} H={tb.H} loops={n:5d}: {ms / E * 1e3:8.1f} us/episode ({n * E * tb.H / ms / 1e6:7.3f} G agent-steps/s), "
              f"of which resample+VI {ms_s / E * 1e3:8.1f} us", flush=True)
        del ag; torch.cuda.empty_cache()
