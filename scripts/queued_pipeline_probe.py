"""End-to-end step of 65,536 envs on the C2 MDP through the two-level host pipeline: colo_env_pipeline_run (stream sync +
launch per group-step) against colo_env_pipeline_run_queued (stream memory operations; one triple at a time, or CUDA-graph
replays), for 2..4 groups, int32 and compact host I/O.  python scripts/queued_pipeline_probe.py [n_steps]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench
from colosseum_b200.batched_mdp import PipelinedBatchedMDP

K = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
N = 65536
tb = bench.load_c2_tables()
gen = torch.Generator().manual_seed(1234)
acts = [torch.randint(0, tb.A, (N,), dtype=torch.int32, generator=gen) for _ in range(8)]
for compact in (False, True):
    for G in (2, 3, 4, 6, 8):
        for runner in ("native", "threads"):
            env = PipelinedBatchedMDP(tb, N, groups=G, mode="dense_f32", seed=1234, compact_io=compact)
            env.reset()
            ring = [[a[o:o + n].to(env.shards[0].action_dtype).clone().pin_memory() for o, n in zip(env.offsets, env.sizes)]
                    for a in acts]
            run = {"native": env.run_native, "threads": lambda r, k: env.run_native(r, k, threads=True),
                   "queued": lambda r, k: env.run_queued(r, k, graph=False),
                   "graph": lambda r, k: env.run_queued(r, k, graph=True)}[runner]
            try:
                run(ring, 200)
                best = 1e9
                for _ in range(3):
                    w0 = time.perf_counter()
                    run(ring, K)
                    best = min(best, time.perf_counter() - w0)
                ok = all(int(sh.status.item()) == 0 for sh in env.shards)
                print(f"compact={int(compact)} groups={G} {runner:7s}: {best / K * 1e6:7.2f} us/step "
                      f"{N * K / best / 1e9:6.3f} G env-steps/s ok={ok}", flush=True)
            except Exception as e:
                print(f"compact={int(compact)} groups={G} {runner:7s}: {type(e).__name__}: {str(e)[:200]}", flush=True)
            del env
