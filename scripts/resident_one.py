"""single resident-solver launches for ncu"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from resident_probe import run
for name, nv in (("doc_simplegrid4", 1), ("c2_deepsea30_prand", 1), ("c2_deepsea30_prand", 4)):
    g = np.load(f"tests/golden/inst_{name}.npz")
    T, R = torch.from_numpy(g["T"]).cuda(), torch.from_numpy(g["R"]).cuda()
    S = T.shape[0]
    if nv == 1:
        print(name, run(T, R, 2000))
    else:
        pins = torch.arange(36, dtype=torch.int32, device="cuda")
        print(name, "nv4", run(T, None, 500, NV=4, pins=pins, fold=2, gamma=1.0, r_const=1.0))
