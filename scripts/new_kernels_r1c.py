"""one invocation each of the kernels added late in round 1 (for ncu): sparse Gauss-Seidel solve, Q-learning loops"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import colosseum_b200.dynamic_programming as dp
import colosseum_b200.agent_loop as al
from colosseum_b200.suite import load_suite
from colosseum_b200.tables import MDPTables
inst = [i for i in load_suite("tests/golden/c3_suite.npz") if i.name.startswith("MiniGridRoomsContinuous.ergo1")][0]
T = torch.from_numpy(inst.tables.T).cuda(); R = torch.from_numpy(inst.R).cuda()
Q, V = dp.discounted_value_iteration(T, R, sweep_order="gauss_seidel")
print("sparse GS VI", inst.S, float(V[0]), dp.last_iterations())
tb = MDPTables.from_golden(np.load("tests/golden/inst_c2_deepsea30_prand.npz"))
ag = al.QLearningContinuous(1, tb, 10 ** 6, n_loops=65536)
ag.steps(100)
torch.cuda.synchronize()
print("agents", float(ag.cumulative_reward.mean()))
tb = MDPTables.from_golden(np.load("tests/golden/inst_taxi_epi.npz"))
ag = al.QLearningEpisodic(1, tb, 10 ** 6, p=0.05, c_1=0.5, c_2=0.5, UCB_type="bernstein", n_loops=16384)
ag.steps(100)
torch.cuda.synchronize()
print("agents epi", float(ag.cumulative_reward.mean()))
