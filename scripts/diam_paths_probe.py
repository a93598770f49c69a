"""GPU probe: continuous diameter on suite instances under the implementation selected by COLO_DIAM_PATH."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import colosseum_b200.hardness as hd
from colosseum_b200.suite import load_suite

names = sys.argv[1:] or ["TaxiContinuous.ergo0", "DeepSeaContinuous.ergo0", "SimpleGridContinuous.ergo2",
                         "MiniGridEmptyContinuous.ergo0", "MiniGridEmptyContinuous.ergo3", "MiniGridRoomsContinuous.ergo1"]
suite = {i.name: i for i in load_suite("tests/golden/c3_suite.npz")}
print("path =", os.environ.get("COLO_DIAM_PATH", "auto"))
for n in names:
    inst = suite[n]
    T = torch.from_numpy(inst.tables.T).cuda()
    for prec in ("f32", "f64"):
        hd.get_diameter(T, False, precision=prec, max_iter=50) if False else None
        torch.cuda.synchronize(); t0 = time.perf_counter()
        d, sw = hd.get_diameter(T, False, precision=prec, return_sweeps=True)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        ref = inst.ref["cached_diameter"]
        print(f"{n:34s} S={inst.S:4d} A={inst.A} {prec} d={d:10.5f} (cached {ref:10.5f}) sweeps={sw:6d} {dt*1e3:9.2f} ms {dt/sw*1e6:8.1f} us/sweep")
