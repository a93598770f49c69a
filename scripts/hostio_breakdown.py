"""Where an end-to-end host_io step spends its time: kernel duration under zero-copy I/O (CUDA events) vs the
device-resident kernel, launch+sync floor, for several batch sizes."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from colosseum_b200.batched_mdp import BatchedMDP
from colosseum_b200 import _cabi
import bench

tb = bench.load_c2_tables()
for N in (8192, 16384, 32768, 65536, 131072):
    rng = np.random.default_rng(0)
    acts_h = [torch.from_numpy(rng.integers(0, tb.A, N).astype(np.int32)).pin_memory() for _ in range(4)]
    acts_d = [a.cuda() for a in acts_h]
    env_h = BatchedMDP(tb, N, host_io=True); env_h.reset()
    env_d = BatchedMDP(tb, N); env_d.reset()
    K = 200
    res = {}
    for name, env, acts in (("zero-copy", env_h, acts_h), ("device", env_d, acts_d)):
        for i in range(20):
            env.step_async(acts[i % 4], auto_reset=True)
        torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        for i in range(K):
            evs[i][0].record()
            env.step_async(acts[i % 4], auto_reset=True)
            evs[i][1].record()
            torch.cuda.synchronize()
        t = np.array([a.elapsed_time(b) for a, b in evs]) * 1e3
        res[name] = np.median(t)
    # wall clock of the lean path
    for i in range(20): env_h.step_host(acts_h[i % 4], auto_reset=True)
    t0 = time.perf_counter()
    for i in range(K): env_h.step_host(acts_h[i % 4], auto_reset=True)
    wall = (time.perf_counter() - t0) / K * 1e6
    # copy-engine transfer of the same bytes
    out_h = torch.empty(9 * N, dtype=torch.uint8).pin_memory()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(50): out_h.copy_(env_d._out, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    d2h = e0.elapsed_time(e1) / 50 * 1e3
    e0.record()
    for i in range(50): env_d.action.copy_(acts_h[0], non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    h2d = e0.elapsed_time(e1) / 50 * 1e3
    print(f"N={N:6d} kernel zero-copy {res['zero-copy']:6.1f} us  device {res['device']:6.1f} us  step_host wall {wall:6.1f} us  "
          f"DMA d2h {d2h:5.1f} us ({9*N/d2h/1e3:.1f} GB/s) h2d {h2d:5.1f} us ({4*N/h2d/1e3:.1f} GB/s)", flush=True)
# launch + sync floor
lib = _cabi.lib()
s = _cabi.current_stream()
x = torch.zeros(1, device="cuda")
t0 = time.perf_counter()
for i in range(1000):
    lib.colo_stream_synchronize(s)
print(f"idle sync call {(time.perf_counter() - t0) / 1000 * 1e6:.2f} us")
env = BatchedMDP(tb, 32, host_io=True); env.reset()
a = torch.zeros(32, dtype=torch.int32).pin_memory()
for i in range(50): env.step_host(a, auto_reset=True)
t0 = time.perf_counter()
for i in range(1000): env.step_host(a, auto_reset=True)
print(f"step_host N=32 wall {(time.perf_counter() - t0) / 1000 * 1e6:.2f} us")
