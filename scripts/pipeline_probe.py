"""e2e step throughput of PipelinedBatchedMDP vs groups (host actions in pinned memory, TimeStep in pinned memory)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from colosseum_b200 import synth
from colosseum_b200.batched_mdp import BatchedMDP, PipelinedBatchedMDP
import bench

tb = bench.load_c2_tables()
N = 65536
for groups in (1, 2, 3, 4, 5, 6):
    env = PipelinedBatchedMDP(tb, N, groups=groups)
    env.reset()
    rng = np.random.default_rng(0)
    acts = [[torch.from_numpy(rng.integers(0, tb.A, env.sizes[g]).astype(np.int32)).pin_memory() for g in range(groups)]
            for _ in range(4)]
    for it in range(20):
        env.step_all(acts[it % 4])
    # pipelined loop
    for mode in ("step_all", "pipelined"):
        K = 300
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if mode == "step_all":
            for it in range(K):
                env.step_all(acts[it % 4])
        else:
            for g in range(groups): env.send(g, acts[0][g])
            for it in range(1, K):
                a = acts[it % 4]
                for g in range(groups):
                    env.recv(g)
                    env.send(g, a[g])
            for g in range(groups): env.recv(g)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"groups={groups} {mode:10s} {dt / K * 1e6:7.2f} us/step  {N * K / dt / 1e9:6.3f} G env-steps/s", flush=True)
