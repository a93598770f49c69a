"""Does one host thread per env group beat the single-thread pipelined loop? (ctypes releases the GIL in launch/sync)"""
import os, sys, time, threading
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from colosseum_b200.batched_mdp import PipelinedBatchedMDP
import bench
tb = bench.load_c2_tables()
N = 65536
for groups in (2, 3, 4):
    env = PipelinedBatchedMDP(tb, N, groups=groups); env.reset()
    rng = np.random.default_rng(0)
    acts = [torch.from_numpy(rng.integers(0, tb.A, env.sizes[g]).astype(np.int32)).pin_memory() for g in range(groups)]
    K = 2000
    def work(g):
        sh = env.shards[g]; a = acts[g]
        for it in range(K):
            sh.send_host(a); sh.recv_host()
    for g in range(groups): env.shards[g].send_host(acts[g]); env.shards[g].recv_host()
    ths = [threading.Thread(target=work, args=(g,)) for g in range(groups)]
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for t in ths: t.start()
    for t in ths: t.join()
    dt = time.perf_counter() - t0
    print(f"groups={groups} one thread per group: {dt / K * 1e6:6.2f} us/step  {N * K / dt / 1e9:.3f} G env-steps/s", flush=True)
