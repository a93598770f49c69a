"""single sparse diameter solves (for ncu / NT sweeps)"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import colosseum_b200.hardness as hd
from colosseum_b200.suite import load_suite
names = ["MiniGridRoomsContinuous.ergo1", "SimpleGridContinuous.comm2", "TaxiContinuous.ergo0"]
suite = {i.name: i for i in load_suite("tests/golden/c3_suite.npz", only=set(names))}
for n in names:
    T = torch.from_numpy(suite[n].tables.T).cuda()
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        d, sw = hd.get_diameter(T, False, precision="f64", return_sweeps=True)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"NT={os.environ.get('COLO_SPARSE_NT','auto')} {n:32s} d={d:.4f} sweeps={sw} {dt*1e3:.2f} ms {dt/sw*1e6:.1f} us/sweep")
