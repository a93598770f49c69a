"""Two-group pipelined end-to-end step with the TimeStep block written to DEVICE memory by the kernel and moved to pinned
host memory by the copy engine (cudaMemcpyAsync on the group's stream), against the shipped zero-copy stores; Python loop
for both.  python scripts/e2e_ce_probe.py"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench
from colosseum_b200.batched_mdp import BatchedMDP

tb = bench.load_c2_tables()
N, K = 65536, 4000
rt = C.CDLL("libcudart.so.12")
rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
rt.cudaStreamSynchronize.argtypes = [C.c_void_p]


def make(n, out_dev, act_dev, stream, off):
    env = BatchedMDP(tb, n, mode="dense_f32", seed=1, host_io=True, compact_io=True, stream=stream, env_offset=off)
    with torch.cuda.stream(stream):
        env.reset()
    x = {"env": env, "s": C.c_void_p(int(stream.cuda_stream)), "n": n}
    if out_dev:
        d = torch.zeros(7 * n, dtype=torch.uint8, device="cuda")
        env._batch.reward, env._batch.obs, env._batch.step_type_mirror = d.data_ptr(), d.data_ptr() + 4 * n, d.data_ptr() + 6 * n
        x["dev_out"] = d
    a = torch.randint(0, tb.A, (n,), dtype=torch.uint8).pin_memory()
    x["a_host"] = a
    x["a_dev"] = torch.zeros(n, dtype=torch.uint8, device="cuda") if act_dev else None
    torch.cuda.synchronize()
    env._make_stepper()
    return x


def step(x):
    env = x["env"]
    if x["a_dev"] is not None:
        rt.cudaMemcpyAsync(x["a_dev"].data_ptr(), x["a_host"].data_ptr(), x["n"], 1, x["s"])
        env.send_host(x["a_dev"])
    else:
        env.send_host(x["a_host"])
    if "dev_out" in x:
        rt.cudaMemcpyAsync(env._out.data_ptr(), x["dev_out"].data_ptr(), 7 * x["n"], 2, x["s"])


for G in (1, 2, 3, 4):
    for out_dev, act_dev in ((False, False), (True, False), (True, True)):
        sizes = [N // G] * G
        gs = [make(sizes[g], out_dev, act_dev, torch.cuda.Stream(), g * (N // G)) for g in range(G)]
        best = 1e9
        for rep in range(3):
            for x in gs:
                step(x)
            w0 = time.perf_counter()
            for i in range(K):
                for x in gs:
                    rt.cudaStreamSynchronize(x["s"])
                    step(x)
            for x in gs:
                rt.cudaStreamSynchronize(x["s"])
            best = min(best, time.perf_counter() - w0)
        print(f"groups={G} outputs {'device + copy engine' if out_dev else 'zero-copy stores   '} actions "
              f"{'copy engine' if act_dev else 'zero-copy  '}: {best / K * 1e6:6.2f} us/step {N * K / best / 1e9:5.2f} G env-steps/s", flush=True)
