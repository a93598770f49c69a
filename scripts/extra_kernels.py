"""one invocation each of the secondary kernels (for ncu): dense multi-target GEMM diameter, resident cluster solver,
extended VI, stationary-distribution squaring, Dirichlet sampling, episodic sparse diameter"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import colosseum_b200.dynamic_programming as dp
import colosseum_b200.hardness as hd
import colosseum_b200.markov_chain as mc
from colosseum_b200.posterior import sample_transition_model
rs = np.random.RandomState(0)
S, A = 768, 4
T = rs.dirichlet(np.ones(S) * 0.05, size=(S, A)).astype(np.float32)   # dense rows -> GEMM path
R = rs.rand(S, A).astype(np.float32)
Td = torch.from_numpy(T).cuda()
print("dense diameter (gemm):", hd.get_diameter(Td, False, precision="f32", epsilon=1e-3))
T2 = rs.dirichlet(np.ones(200) * 0.3, size=(200, 3)).astype(np.float32)  # small dense -> resident cluster solver
print("resident VI:", float(dp.discounted_value_iteration(T2, rs.rand(200, 3).astype(np.float32), 0.99, 1e-4)[1][0]))
g = np.load("tests/golden/evi.npz")
print("EVI span:", dp.extended_value_iteration(g["P_3"], g["est_3"], g["beta_r_3"], g["beta_p_3"], 1.0, 1e-3)[0])
pi = np.full((S, A), 1.0 / A, np.float32)
print("avg reward:", mc.get_average_reward(Td, R, pi))
print("dirichlet:", float(sample_transition_model(torch.full((S, A, S), 0.3, device="cuda"), seed=1).sum()))
ge = np.load("tests/golden/inst_taxi_epi.npz")
print("episodic diameter:", hd.get_diameter(ge["T_epi"], True))
