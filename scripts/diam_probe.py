"""GPU probe: diameter / VI solve latency on the golden instances."""
import glob, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import colosseum_b200.hardness as hd
import colosseum_b200.dynamic_programming as dp

for name in ["doc_simplegrid4", "frozenlakecontinuous_ergo0", "taxicontinuous_ergo0", "simplegridcontinuous_ergo1",
             "deepsea20_prand", "c2_deepsea30_prand"]:
    g = np.load(f"tests/golden/inst_{name}.npz")
    T, R = torch.from_numpy(g["T"]).cuda(), torch.from_numpy(g["R"]).cuda()
    S, A = R.shape
    for prec in ("f64", "f32"):
        hd.get_diameter(T, False, precision=prec)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        d, sw = hd.get_diameter(T, False, precision=prec, return_sweeps=True)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"{name:28s} S={S:4d} A={A} {prec} diameter={d:10.4f} sweeps={sw:6d} {dt*1e3:9.2f} ms  {dt/sw*1e6:7.1f} us/sweep")
    torch.cuda.synchronize(); t0 = time.perf_counter()
    Q, V = dp.discounted_value_iteration(T, R, 0.99, 1e-3)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    it = dp.last_iterations()[0]
    print(f"{'':28s} VI eps=1e-3: {it} sweeps {dt*1e3:.2f} ms  {dt/it*1e6:.1f} us/sweep")
for name in ["taxi_epi", "minigridempty5_epi"]:
    g = np.load(f"tests/golden/inst_{name}.npz")
    T = torch.from_numpy(g["T_epi"]).cuda()
    hd.get_diameter(T, True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    d, sw = hd.get_diameter(T, True, return_sweeps=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"{name:28s} T_epi={tuple(T.shape)} diameter={d:.4f} iters={sw} {dt*1e3:.2f} ms")
