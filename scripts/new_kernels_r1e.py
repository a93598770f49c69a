"""one PSRL episode boundary (for ncu): Dirichlet sample (fast), NIG sample, batched episodic VI, PSRL steps"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import colosseum_b200.agent_loop as al
from colosseum_b200.tables import MDPTables
tb = MDPTables.from_golden(np.load("tests/golden/inst_taxi_epi.npz"))
ag = al.PSRLEpisodic(0, tb, 10 ** 6, n_loops=1024)
ag.steps(2 * tb.H)
torch.cuda.synchronize()
print("psrl", float(ag.cumulative_reward.mean()))
