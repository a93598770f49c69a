"""Where the C3 hardness phase spends its time: per-phase totals over the 80 suite instances (one at a time)."""
import os, sys, time, collections
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from colosseum_b200.suite import load_suite, hardness_of_instance
insts = list(load_suite("tests/golden/c3_suite.npz"))
for inst in insts[:6]:
    hardness_of_instance(inst)  # warm-up
tot = collections.Counter(); per = []
for inst in insts:
    tm = {}
    t0 = time.perf_counter()
    hardness_of_instance(inst, timings=tm)
    per.append((time.perf_counter() - t0, inst.name, inst.S, inst.tables.H, tm))
    tot.update(tm)
print({k: round(v, 3) for k, v in tot.most_common()}, "total", round(sum(tot.values()), 3))
for dt, name, S, H, tm in sorted(per, reverse=True)[:12]:
    print(f"{dt*1e3:7.1f} ms {name:34s} S={S} H={H}", {k: round(v * 1e3, 1) for k, v in sorted(tm.items(), key=lambda kv: -kv[1])[:4]})
