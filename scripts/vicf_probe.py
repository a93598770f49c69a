"""GPU probe: discounted VI on the continuous form of TaxiEpisodic.comm1 (n ~ 6000 nodes) and on MiniGridRooms S=948"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import colosseum_b200.dynamic_programming as dp
import colosseum_b200.episodic_forms as ef
from colosseum_b200.suite import load_suite
suite = {i.name: i for i in load_suite("tests/golden/c3_suite.npz", only={"TaxiEpisodic.comm1", "MiniGridRoomsContinuous.ergo1"})}
inst = suite["TaxiEpisodic.comm1"]
tb = inst.tables
T = torch.from_numpy(tb.T).cuda(); R = torch.from_numpy(inst.R).cuda()
sp = np.diff(tb.start_cum, prepend=0.0)
T_cf, R_cf = ef.get_continuous_form_episodic_transition_matrix_and_rewards(tb.H, T, R, tb.start_idx, sp, nodes=inst.nodes)
print("T_cf", tuple(T_cf.shape), "max nnz/row", int((T_cf != 0).sum(-1).max()))
i2 = suite["MiniGridRoomsContinuous.ergo1"]
T2 = torch.from_numpy(i2.tables.T).cuda(); R2 = torch.from_numpy(i2.R).cuda()
for name, (TT, RR) in {"taxi_cf": (T_cf, R_cf), "rooms948": (T2, R2)}.items():
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        Q, V = dp.discounted_value_iteration(TT, RR, 0.99, 1e-9, precision="f64")
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"{name} C={os.environ.get('COLO_SPARSE_C','auto')} nosparse={os.environ.get('COLO_NO_SPARSE')} rep{rep}: {dt*1e3:.2f} ms iters={dp.last_iterations()[0]} V0={float(V[0]):.6f}")
