"""Small fixed launch sequences for the round-2 ncu captures (scripts/profile_r2.sh): one target per invocation.

    python scripts/ncu_targets_r2.py step N | backup_c4 | backup_c5 | gs | umma | epi_batched | hitting_gemm
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import colosseum_b200.dynamic_programming as dp  # noqa: E402
from colosseum_b200 import _cabi  # noqa: E402

what = sys.argv[1]
if what == "step":
    from colosseum_b200.batched_mdp import BatchedMDP

    N = int(sys.argv[2])
    tb = bench.load_c2_tables()
    env = BatchedMDP(tb, N, mode="dense_f32", seed=5)
    env.reset()
    act = torch.randint(0, tb.A, (N,), dtype=torch.int32, device="cuda")
    for _ in range(5):
        env.step_async(act, auto_reset=True)
elif what in ("backup_c4", "backup_c5"):
    if what == "backup_c4":  # the bench shape: 4,096 MDPs S=512 A=4 (17.2 GB)
        T, R = bench.make_c4_batch(4096, 512, 4, seed=100)
    else:  # the C5 kernel on one MDP S=16,384 A=8 (8.6 GB: what ncu's save / restore handles in minutes; bench: 40,000)
        from colosseum_b200.synth import synth_dense_rows

        S5 = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
        T, R = synth_dense_rows(0, S5, S5, 8, seed=7)
    vi = dp.BatchedValueIteration(T, R, gamma=0.99, precision="f32")
    vi.sweep(5)
elif what == "gs":
    T, R = bench.make_c4_batch(1036, 512, 4, seed=100)  # one full round of the 148 x 7 warps
    dp.discounted_value_iteration(T, R, 0.99, 1e-3, sweep_order="gauss_seidel")
elif what == "umma":
    g = torch.Generator(device="cuda").manual_seed(0)
    S, A = 2048, 8
    x = torch._standard_gamma(torch.full((S, A, S), 0.05, device="cuda"), generator=g) + 1e-30
    T = (x / x.sum(-1, keepdim=True)).float().contiguous()
    tg = torch.arange(S, dtype=torch.int32, device="cuda")
    E = torch.zeros((S, S), dtype=torch.float32, device="cuda")
    W = torch.empty_like(E)
    rc = _cabi.lib().colo_hitting_umma_sweeps_f32(_cabi.ptr(T), _cabi.ptr(tg), S, S, A, 4, 1, _cabi.ptr(E), _cabi.ptr(W),
                                                  _cabi.current_stream())
    _cabi.check(rc, "umma")
elif what == "epi_batched":
    rs = np.random.RandomState(0)
    B, S, A, H = 1024, 108, 6, 24
    T = torch.from_numpy(rs.dirichlet(np.ones(S) * 0.1, size=(B, S, A)).astype(np.float32)).cuda()
    R = torch.rand((B, S, A), device="cuda")
    for _ in range(3):
        dp.episodic_value_iteration(H, T, R, precision="f32")
torch.cuda.synchronize()
print("ok", what)
