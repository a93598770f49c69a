"""Host-side cost of the pipelined loop: average duration of send() and recv() calls per group count."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from colosseum_b200.batched_mdp import PipelinedBatchedMDP
import bench
tb = bench.load_c2_tables()
N = 65536
pc = time.perf_counter
for groups in (2, 3, 4):
    env = PipelinedBatchedMDP(tb, N, groups=groups); env.reset()
    rng = np.random.default_rng(0)
    acts = [torch.from_numpy(rng.integers(0, tb.A, env.sizes[g]).astype(np.int32)).pin_memory() for g in range(groups)]
    for g in range(groups): env.send(g, acts[g])
    for it in range(50):
        for g in range(groups): env.recv(g); env.send(g, acts[g])
    K = 400; ts = tr = 0.0
    t0 = pc()
    for it in range(K):
        for g in range(groups):
            a = pc(); env.recv(g); b = pc(); env.send(g, acts[g]); c = pc()
            tr += b - a; ts += c - b
    tot = pc() - t0
    for g in range(groups): env.recv(g)
    print(f"groups={groups}: {tot / K * 1e6:6.2f} us/step; per group-step recv {tr / K / groups * 1e6:5.2f} us, send {ts / K / groups * 1e6:5.2f} us", flush=True)
# empty-stream costs
from colosseum_b200 import _cabi
lib = _cabi.lib(); s = torch.cuda.Stream(); sp = int(s.cuda_stream)
t0 = pc()
for i in range(2000): lib.colo_stream_synchronize(sp)
print(f"sync on an idle stream: {(pc() - t0) / 2000 * 1e6:.2f} us")
