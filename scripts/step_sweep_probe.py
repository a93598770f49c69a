"""Env-count sweep of the step kernel with attribution switches (visitation counters on/off, kernel variant), and the
small fixed launch sequence that scripts/profile_r2.sh captures under ncu.

    python scripts/step_sweep_probe.py                 # timing table (CUDA events, graph of 8 steps, >= 20 ms)
    python scripts/step_sweep_probe.py --ncu N         # 3 warm + 2 plain launches at N envs (for ncu -k regex:env_step)
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import load_c2_tables  # noqa: E402
from colosseum_b200.batched_mdp import BatchedMDP  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--ncu", type=int, default=0)
ap.add_argument("--mode", default="dense_f32")
args = ap.parse_args()
tb = load_c2_tables()
gen = torch.Generator(device="cuda").manual_seed(1)

if args.ncu:
    N = args.ncu
    env = BatchedMDP(tb, N, mode=args.mode, seed=5)
    env.reset()
    act = torch.randint(0, tb.A, (N,), dtype=torch.int32, device="cuda", generator=gen)
    for _ in range(5):
        env.step_async(act, auto_reset=True)
    torch.cuda.synchronize()
    print("ok", N)
    sys.exit(0)

rows = []
for N in (65536, 1 << 20, 1 << 22, 1 << 24):
    for visits in (True, False):
        for mode in ("dense_f32", "succ"):
            env = BatchedMDP(tb, N, mode=mode, seed=5, track_visits=visits)
            env.reset()
            act = torch.randint(0, tb.A, (N,), dtype=torch.int32, device="cuda", generator=gen)
            for _ in range(3):
                env.step_async(act, auto_reset=True)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _ in range(8):
                    env.step_async(act, auto_reset=True)
            g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record(); e1.synchronize()
            reps = max(1, int(20.0 / max(e0.elapsed_time(e1), 1e-3)))
            e0.record()
            for _ in range(reps):
                g.replay()
            e1.record(); e1.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / (8 * reps)
            rows.append(dict(n_envs=N, visits=visits, mode=mode, us_per_step=us, env_steps_per_s=N / us * 1e6,
                             state_stream_gbs=34 * N / us / 1e3))
            print(rows[-1], flush=True)
            del env, g, act
            torch.cuda.empty_cache()
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "r2_step_sweep_probe.json"), "w"), indent=1)
