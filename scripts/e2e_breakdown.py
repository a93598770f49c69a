"""GPU probe: where the end-to-end step time goes (host_io mode)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from colosseum_b200.batched_mdp import BatchedMDP
from colosseum_b200.tables import MDPTables

g = np.load("tests/golden/inst_c2_deepsea30_prand.npz")
tb = MDPTables.from_golden(g)
N = 65536
gen = torch.Generator().manual_seed(0)
h_act = [torch.randint(0, tb.A, (N,), dtype=torch.int32, generator=gen).pin_memory() for _ in range(8)]
d_act = [a.cuda() for a in h_act]
envh = BatchedMDP(tb, N, mode="dense_f32", seed=1234, host_io=True); envh.reset()
env = BatchedMDP(tb, N, mode="dense_f32", seed=1234); env.reset()

def kernel_us(e, acts, n=200):
    for i in range(20): e.step_async(acts[i % 8], auto_reset=True)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for i, (a, b) in enumerate(ev):
        a.record(); e.step_async(acts[i % 8], auto_reset=True); b.record()
        torch.cuda.synchronize()
    return np.median([a.elapsed_time(b) for a, b in ev]) * 1e3

print(f"kernel, device I/O            : {kernel_us(env, d_act):.1f} us")
print(f"kernel, host actions + host out: {kernel_us(envh, h_act):.1f} us")
print(f"kernel, dev actions + host out : {kernel_us(envh, d_act):.1f} us")
# python launch cost: back-to-back launches without sync
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(2000): env.step_async(d_act[i % 8], auto_reset=True)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"python step_async issue cost   : {(t1 - t0) / 2000 * 1e6:.1f} us/launch (drain {(t2 - t1) * 1e3:.2f} ms)")
# sync latency on an empty stream
t0 = time.perf_counter()
for i in range(2000): torch.cuda.current_stream().synchronize()
print(f"idle stream sync               : {(time.perf_counter() - t0) / 2000 * 1e6:.2f} us")
# launch + sync of the device-I/O kernel
t0 = time.perf_counter()
for i in range(1000):
    env.step_async(d_act[i % 8], auto_reset=True); torch.cuda.current_stream().synchronize()
print(f"device-I/O step + sync         : {(time.perf_counter() - t0) / 1000 * 1e6:.1f} us")
t0 = time.perf_counter()
for i in range(1000):
    envh.step_async(h_act[i % 8], auto_reset=True); torch.cuda.current_stream().synchronize()
print(f"host-I/O step + sync           : {(time.perf_counter() - t0) / 1000 * 1e6:.1f} us")
