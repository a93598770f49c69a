import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import colosseum_b200.dynamic_programming as dp
from bench import make_c4_batch
for B in (148 * 8, 4096):
    T, R = make_c4_batch(B, 512, 4, seed=3)
    for order in ("jacobi", "gauss_seidel"):
        dp.discounted_value_iteration(T[:8], R[:8], 0.99, 1e-3, sweep_order=order)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        Q, V = dp.discounted_value_iteration(T, R, 0.99, 1e-3, sweep_order=order)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        it = np.asarray(dp.last_iterations())
        gb = it.mean() * 4 * 512 * 4 * 512 * B / dt / 1e9
        print(f"B={B} {order:13s}: {dt*1e3:8.1f} ms  sweeps mean {it.mean():.0f} max {it.max()}  -> {gb:7.0f} GB/s of T")
    del T, R
