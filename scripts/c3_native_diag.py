"""Diagnostics of the native C3 runner: parity failures against the recorded reference answers and the slowest
instances, for a given worker count (argv[1], default 8)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from colosseum_b200.suite import load_suite_all, run_many_native
W = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n = int(sys.argv[2]) if len(sys.argv) > 2 else 160
g = os.path.join(ROOT, "tests", "golden")
suite = load_suite_all(g, indices=list(range(80)))
work = [(suite[i % 80], 0) for i in range(n)]
run_many_native(work[:2], n_workers=2, n_envs=1024, n_steps=10)
for rep in range(2):
    t0 = time.perf_counter()
    res = run_many_native(work, n_workers=W, n_envs=1024, n_steps=1000)
    dt = time.perf_counter() - t0
    print(f"rep {rep}: {n / dt:.1f} inst/s with {W} workers", flush=True)
per = sorted(((tm["hardness_s"] + tm["step_s"], inst.name, tm) for (inst, _), (r, tm) in zip(work, res)), reverse=True)
for t, name, tm in per[:6]:
    print(f"{t*1e3:8.1f} ms {name} step {tm['step_s']*1e3:.1f} hard {tm['hardness_s']*1e3:.1f}")
bad = 0
for (inst, _), (r, tm) in list(zip(work, res))[:80]:
    for k, rk, tol in (("diameter", "diameter", 2e-3), ("diameter", "cached_diameter", 2e-3), ("value_norm", "value_norm", 3e-3),
                       ("value_norm", "cached_value_norm", 3e-3), ("gaps", "gaps", 7e-3)):
        ref = inst.ref.get(rk, float("nan"))
        if ref == ref and r[k] == r[k] and abs(r[k] - ref) > tol * max(abs(ref), 1e-3):
            bad += 1
            print("PARITY", inst.name, rk, r[k], ref, abs(r[k] - ref) / max(abs(ref), 1e-3))
print("parity failures:", bad)
