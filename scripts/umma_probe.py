"""Throughput / accuracy probe of the tcgen05 hitting-time sweep (hitting_umma.cu): one process, COLO_UMMA_FLUSH
(k-blocks per TMEM accumulator chain; read at every plan) varied in a loop.  Results are appended to
gpurun_out/r2_umma_probe.jsonl as they come."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from colosseum_b200 import _cabi  # noqa: E402
from colosseum_b200.suite import load_suite  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out", "r2_umma_probe.jsonl")
T0 = time.perf_counter()


def emit(d):
    d["t"] = round(time.perf_counter() - T0, 1)
    with open(OUT, "a") as f:
        f.write(json.dumps(d) + "\n")
    print(d, flush=True)


def run(T, targets, n):
    S, A, _ = T.shape
    K = len(targets)
    E = torch.zeros((K, S), dtype=torch.float32, device="cuda")
    W = torch.empty_like(E)
    rc = _cabi.lib().colo_hitting_umma_sweeps_f32(_cabi.ptr(T), _cabi.ptr(targets), K, S, A, n, 1, _cabi.ptr(E),
                                                  _cabi.ptr(W), _cabi.current_stream())
    _cabi.check(rc, "umma")
    return E


def timed(T, targets, n0, n1):
    ts = []
    for n in (n0, n1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run(T, targets, n)
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return (ts[1] - ts[0]) / (n1 - n0) * 1e3  # us per sweep


def fp64_sweeps(T, targets, n):
    S, A = T.shape[0], T.shape[1]
    Td = T.double().reshape(-1, S)
    K = len(targets)
    E = torch.zeros((K, S), dtype=torch.float64, device="cuda")
    idx = torch.arange(K, device="cuda")
    for _ in range(n):
        E = (1.0 + (Td @ E.T).reshape(S, A, K)).min(1).values.T.contiguous()
        E[idx, targets.long()] = 0.0
    return E


torch.cuda.set_per_process_memory_fraction(0.5)
pool = torch.cuda.current_device()
suite = load_suite(os.path.join(ROOT, "tests", "golden", "c3_suite.npz"))
inst = max((i for i in suite if not i.episodic), key=lambda i: i.S)
cases = {"rooms948": torch.from_numpy(inst.tables.T).cuda()}
g = torch.Generator(device="cuda").manual_seed(0)
for name, (S, A) in {"dense1024x4": (1024, 4), "dense2048x8": (2048, 8)}.items():
    x = torch._standard_gamma(torch.full((S, A, S), 0.05, device="cuda"), generator=g) + 1e-30
    cases[name] = (x / x.sum(-1, keepdim=True)).float().contiguous()
emit({"event": "setup done"})
N_ACC = 20
refs = {}
for name, T in cases.items():
    targets = torch.arange(T.shape[0], dtype=torch.int32, device="cuda")
    refs[name] = fp64_sweeps(T, targets, N_ACC)
torch.cuda.synchronize()
emit({"event": "fp64 references done"})
for flush, bn, mc in ((1, 0, 1), (1, 0, 0), (2, 0, 1), (1, 128, 1), (1, 64, 1)):
    os.environ["COLO_UMMA_FLUSH"] = str(flush)
    os.environ["COLO_UMMA_BN"] = str(bn)
    os.environ["COLO_UMMA_CLUSTER"] = str(mc)
    for name, T in cases.items():
        S, A = T.shape[0], T.shape[1]
        targets = torch.arange(S, dtype=torch.int32, device="cuda")
        run(T, targets, 2)
        torch.cuda.synchronize()
        us = min(timed(T, targets, 10, 1010 if S < 1100 else 210) for _ in range(2))
        flop = 2.0 * S * A * S * S
        got = run(T, targets, N_ACC).double()
        ref = refs[name]
        emit({"flush": flush, "bn": bn, "cluster": mc, "case": name, "S": S, "A": A, "us_per_sweep": us, "tflops_fp32_equiv": flop / us / 1e6,
              "tensor_tflops_3x": 3 * flop / us / 1e6,
              f"rel_err_after_{N_ACC}_sweeps": float((got - ref).abs().max() / ref.abs().max()),
              "mean_signed_rel": float((got - ref).sum() / ref.abs().sum())})
