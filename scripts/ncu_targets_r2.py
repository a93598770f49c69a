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
elif what in ("evi_batched", "psrlc_sample", "avg_rewards"):
    # the model-based continuous agents (round 2b): 1,024 FrozenLake loops a few thousand steps in, then ONE planning round
    # for all loops at once
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import load_instance
    import colosseum_b200.agent_loop as al
    from colosseum_b200.tables import MDPTables

    gi = load_instance("frozenlakecontinuous_ergo0")
    tb = MDPTables.from_golden(gi)
    N = 1024
    idx = torch.arange(N, dtype=torch.int32, device="cuda")
    if what == "evi_batched":
        ag = al.UCRL2Continuous(0, tb, 20001, alpha_r=0.1, alpha_p=0.05, bound_type_p="bernstein", n_loops=N)
        ag.steps(3000)
        br, bp = ag.bounds(idx)
        Q, V = torch.zeros_like(ag.Q), torch.zeros_like(ag.V)
        torch.cuda.profiler.start()  # ncu --profile-from-start off: only the all-loops planning round is captured
        for _ in range(2):
            ag.solve_optimistic_model(idx, br, bp, Q, V)
    elif what == "psrlc_sample":
        ag = al.PSRLContinuous(0, tb, 20001, psi_weight=0.02, eta_weight=1e-8, n_loops=N)
        ag.steps(3000)
        torch.cuda.profiler.start()
        for _ in range(2):
            ag.sample_models(idx)
    else:
        import colosseum_b200.markov_chain as mc

        ag = al.QLearningContinuous(0, tb, 20000, n_loops=N)
        ag.steps(3000)
        pi = torch.nn.functional.one_hot(ag.Q.argmax(-1), tb.A).float().contiguous()
        torch.cuda.profiler.start()
        mc.get_average_reward_batched(np.asarray(gi["T"], np.float32), np.asarray(gi["R"], np.float32), pi,
                                      ag.state.cpu().numpy())
torch.cuda.synchronize()
print("ok", what)
