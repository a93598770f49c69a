"""Latency floor of one tiny step: device-resident vs zero-copy host I/O, default stream vs own stream."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from colosseum_b200.batched_mdp import BatchedMDP
from colosseum_b200 import _cabi
import bench
tb = bench.load_c2_tables()
lib = _cabi.lib()

def wall(f, n=2000):
    for i in range(100): f()
    t0 = time.perf_counter()
    for i in range(n): f()
    return (time.perf_counter() - t0) / n * 1e6

for N in (32, 4096):
    a_h = torch.zeros(N, dtype=torch.int32).pin_memory()
    a_d = a_h.cuda()
    env_d = BatchedMDP(tb, N); env_d.reset()
    s0 = _cabi.current_stream()
    def dev_step():
        env_d.step_async(a_d, auto_reset=True); lib.colo_stream_synchronize(s0)
    print(f"N={N} device step + sync (default stream): {wall(dev_step):.2f} us")
    env_h = BatchedMDP(tb, N, host_io=True); env_h.reset()
    print(f"N={N} step_host (default stream): {wall(lambda: env_h.step_host(a_h, auto_reset=True)):.2f} us")
    st = torch.cuda.Stream()
    env_s = BatchedMDP(tb, N, host_io=True, stream=st)
    with torch.cuda.stream(st): env_s.reset()
    torch.cuda.synchronize()
    print(f"N={N} step_host (own stream): {wall(lambda: env_s.step_host(a_h, auto_reset=True)):.2f} us")
    # host_io but actions on the device (only the writes cross PCIe)
    print(f"N={N} step_host-like, device actions, host outputs: "
          f"{wall(lambda: (env_h.step_async(a_d, auto_reset=True), lib.colo_stream_synchronize(s0))):.2f} us")
    env_nv = BatchedMDP(tb, N, host_io=True, track_visits=False); env_nv.reset()
    print(f"N={N} step_host no visit counters: {wall(lambda: env_nv.step_host(a_h, auto_reset=True)):.2f} us")
x = torch.zeros(8, device="cuda")
print(f"torch fill + sync: {wall(lambda: (x.fill_(1.0), torch.cuda.synchronize())):.2f} us")
