"""GPU probe: hardness of every C3 suite instance vs the reference's recorded values, with timings."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from colosseum_b200.suite import load_suite, hardness_of_instance

suite = load_suite("tests/golden/c3_suite.npz")
prec = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else "f64"
rel = lambda a, b: abs(a - b) / max(abs(b), 1e-12)
hardness_of_instance(suite[1], precision=prec)
tot = 0.0
for inst in suite:
    torch.cuda.synchronize(); t0 = time.perf_counter()
    try:
        tm = {}
        res = hardness_of_instance(inst, precision=prec, timings=tm if "--lap" in sys.argv else None)
    except Exception as e:
        print(f"{inst.name:36s} ERROR {type(e).__name__}: {e}"); continue
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    tot += dt
    r = inst.ref
    dref = r["diameter"] if np.isfinite(r["diameter"]) else r["cached_diameter"]
    print(f"{inst.name:36s} S={inst.S:4d} A={inst.A} H={inst.H:3d} {dt*1e3:9.1f} ms  diam {res['diameter']:10.4f} ref {dref:10.4f} "
          f"rel {rel(res['diameter'], dref) if np.isfinite(dref) else float('nan'):.1e} sw {res['diameter_sweeps']:5d} | "
          f"vn {res['value_norm']:.5f} ref {r['value_norm']:.5f} | gaps rel {rel(res['gaps'], r['gaps']):.1e}")
    if "--lap" in sys.argv: print("      " + " ".join(f"{k}={v*1e3:.1f}" for k, v in tm.items()))
print(f"total {tot:.2f} s for {len(suite)} instances ({prec})")
