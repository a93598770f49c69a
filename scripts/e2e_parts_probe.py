"""Where the ~21 us of an end-to-end step go: the compact host_io step kernel launched back to back (no host sync between
launches) with its actions / TimeStep outputs in pinned host memory or in device memory.
python scripts/e2e_parts_probe.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench
from colosseum_b200 import _cabi
from colosseum_b200.batched_mdp import BatchedMDP

tb = bench.load_c2_tables()
K = 2000


def make(N, act_host, out_host, stream):
    env = BatchedMDP(tb, N, mode="dense_f32", seed=1, host_io=True, compact_io=True, stream=stream)
    with torch.cuda.stream(stream):
        env.reset()
    keep = []
    if not out_host:
        dev_out = torch.zeros(7 * N, dtype=torch.uint8, device="cuda")
        keep.append(dev_out)
        env._batch.reward = dev_out.data_ptr()
        env._batch.obs = dev_out.data_ptr() + 4 * N
        env._batch.step_type_mirror = dev_out.data_ptr() + 6 * N
    a = torch.randint(0, tb.A, (N,), dtype=torch.uint8)
    a = a.pin_memory() if act_host else a.cuda()
    keep.append(a)
    torch.cuda.synchronize()
    env._make_stepper()
    return env, a, keep


def run(envs):
    for _ in range(50):
        for env, a, _k in envs:
            env.send_host(a)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for env, a, _k in envs:
        env.stream.wait_event(e0)
    for _ in range(K):
        for env, a, _k in envs:
            env.send_host(a)
    for env, a, _k in envs:
        ev = torch.cuda.Event()
        ev.record(env.stream)
        torch.cuda.current_stream().wait_event(ev)
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) * 1e3 / K


for N in (65536, 32768, 16384):
    for act_host, out_host in ((True, True), (False, True), (True, False), (False, False)):
        envs = [make(N, act_host, out_host, torch.cuda.Stream())]
        us = run(envs)
        print(f"one stream,  {N:6d} envs per launch, actions {'host' if act_host else 'dev '}, outputs {'host' if out_host else 'dev '}: "
              f"{us:6.2f} us per launch", flush=True)
for G in (2, 4):
    for act_host, out_host in ((True, True), (False, True), (True, False)):
        envs = [make(65536 // G, act_host, out_host, torch.cuda.Stream()) for _ in range(G)]
        us = run(envs)
        print(f"{G} streams x {65536 // G:6d} envs, actions {'host' if act_host else 'dev '}, outputs {'host' if out_host else 'dev '}: "
              f"{us:6.2f} us per step of 65,536 envs", flush=True)
