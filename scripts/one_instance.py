"""hardness_of_instance on named suite instances with stage timings (for ncu launch lists)"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from colosseum_b200.suite import load_suite, hardness_of_instance
names = sys.argv[1:] or ["MiniGridRoomsContinuous.ergo1", "TaxiEpisodic.comm1"]
suite = {i.name: i for i in load_suite("tests/golden/c3_suite.npz", only=set(names))}
for n in names:
    for rep in range(2):
        tm = {}
        res = hardness_of_instance(suite[n], precision="f64", timings=tm)
    print(n, {k: round(v * 1e3, 2) for k, v in tm.items()}, res)
