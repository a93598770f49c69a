"""GPU probe: end-to-end step latency, copy-engine path vs zero-copy host I/O (pinned buffers read/written by the kernel)."""
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

def timeit(fn, n=300, w=20):
    for i in range(w): fn(i)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(n): fn(i)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e6

env = BatchedMDP(tb, N, mode="dense_f32", seed=1234); env.reset()
h_out = torch.empty(9 * N, dtype=torch.uint8).pin_memory()
def copy_step(i):
    env.step_async(h_act[i % 8], auto_reset=True); env.fetch_async(h_out); torch.cuda.current_stream().synchronize()
print(f"copy-engine e2e: {timeit(copy_step):.1f} us/step")

envh = BatchedMDP(tb, N, mode="dense_f32", seed=1234, host_io=True); envh.reset()
def host_step(i):
    envh.step_host(h_act[i % 8], auto_reset=True)
print(f"zero-copy  e2e: {timeit(host_step):.1f} us/step")

# hybrid: actions by copy engine, outputs zero-copy
d_act = torch.zeros(N, dtype=torch.int32, device="cuda")
def hybrid_step(i):
    d_act.copy_(h_act[i % 8], non_blocking=True); envh.step_async(d_act, auto_reset=True); torch.cuda.current_stream().synchronize()
print(f"hybrid (H2D copy + zero-copy out): {timeit(hybrid_step):.1f} us/step")
# hybrid 2: zero-copy actions in, outputs by copy engine
def hybrid2_step(i):
    env._batch.action = h_act[i % 8].data_ptr()
    env.step_async(None if False else h_act[i % 8], auto_reset=True); env.fetch_async(h_out); torch.cuda.current_stream().synchronize()
# identical results?
e1 = BatchedMDP(tb, 4096, mode="dense_f32", seed=5); e1.reset()
e2 = BatchedMDP(tb, 4096, mode="dense_f32", seed=5, host_io=True); e2.reset()
a = torch.randint(0, tb.A, (4096,), dtype=torch.int32).pin_memory()
for t in range(50):
    e1.step_async(a, auto_reset=True); o, r, st = e2.step_host(a, auto_reset=True)
torch.cuda.synchronize()
print("same:", torch.equal(e1.obs.cpu(), o), torch.equal(e1.reward.cpu().nan_to_num(7), r.nan_to_num(7)), torch.equal(e1.step_type.cpu(), st))
