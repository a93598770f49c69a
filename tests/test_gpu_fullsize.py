"""GPU: BASELINE.json's configurations at FULL size, through size-independent properties (the oracle cannot run them).

C4: 4,096 synthetic Dirichlet(0.05) MDPs, S=512, A=4 (17.2 GB of T) -- solve to the reference's epsilon, then the
    Bellman residual of every returned V, value bounds, per-instance independence on a sample.
C5: one dense MDP S=40,000, A=8 (51.2 GB of T) -- contraction of successive sweeps, value bounds, and row-shard
    independence: any row range swept alone equals the same rows of the full sweep bit for bit.
Edge cases: empty env batch, single-state MDP, ragged last warp."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GAM = float(np.float32(0.99))


def test_c4_full_size_solve_properties():
    import torch

    import colosseum_b200.dynamic_programming as dp
    from bench import make_c4_batch

    B, S, A = 4096, 512, 4
    free, _ = torch.cuda.mem_get_info()
    if free < 24 << 30:
        pytest.skip("needs 24 GB of free HBM")
    T, R = make_c4_batch(B, S, A, seed=100)
    assert bool(torch.allclose(T.sum(-1), torch.ones((), device="cuda"), atol=1e-5))
    Q, V = dp.discounted_value_iteration(T, R, 0.99, 1e-3)
    iters = np.asarray(dp.last_iterations())
    assert Q.shape == (B, S, A) and V.shape == (B, S) and iters.min() > 100 and iters.max() < 2000
    # fixed-point residual of every instance: one more backup moves V by less than epsilon (the stopping rule)
    Q2, V2 = dp.bellman_backup(T, R, V, gamma=GAM)
    assert float((V2 - V).abs().max()) < 1e-3
    assert bool(torch.equal(Q2.max(-1).values, V2))
    # rewards in [0,1): 0 <= V <= 1/(1-gamma); early-stopped from below (V0 = 0, monotone operator)
    assert float(V.min()) >= 0.0 and float(V.max()) <= 1.0 / (1.0 - GAM) and bool((V2 >= V - 1e-6).all())
    # no cross-instance coupling: instances solved alone give the same bits
    for b in (0, 1777, 4095):
        Qb, Vb = dp.discounted_value_iteration(T[b], R[b], 0.99, 1e-3)
        assert bool(torch.equal(Vb, V[b])) and bool(torch.equal(Qb, Q[b]))


def test_c5_full_size_sweep_properties():
    import torch

    from colosseum_b200.dynamic_programming import BatchedValueIteration
    from colosseum_b200.synth import synth_dense_rows

    S, A = 40000, 8
    free, _ = torch.cuda.mem_get_info()
    if free < 60 << 30:
        pytest.skip("needs 60 GB of free HBM")
    T, R = synth_dense_rows(0, S, S, A, seed=7)
    vi = BatchedValueIteration(T, R, gamma=0.99, precision="f32")
    prev, d_prev = None, None
    for k in range(6):
        vi.sweep(1)
        V = vi.values[0].clone()
        if prev is not None:
            d = float((V - prev).abs().max())
            if d_prev is not None:
                assert d <= GAM * d_prev * (1 + 1e-4), (k, d, d_prev)  # gamma-contraction in the sup norm
            d_prev = d
        prev = V
    assert float(V.min()) >= 0.0 and float(V.max()) <= 6.0 + 1e-3  # six sweeps of rewards in [0,1)
    # row-shard independence (what the multi-GPU split relies on): rows [r0, r0+n) swept alone == rows of the full sweep.
    # Shards of >= 148*64 rows use the same warp-per-state mapping as the full sweep: bit-identical.  Smaller shards
    # (S/8 = 5,000 rows on 8 GPUs) use one CTA per state -- another summation order inside a row: equal to rounding.
    V_in = vi.V[1 - vi.cur].clone()  # the input of the last sweep
    for r0, n in ((20000, 12000), (0, 5000), (17001, 3333), (39990, 10)):
        part = BatchedValueIteration(T[r0:r0 + n], R[r0:r0 + n], gamma=0.99, precision="f32", row0=r0, S_total=S)
        part.V[part.cur].copy_(V_in)
        part.sweep(1)
        if n >= 148 * 64:
            assert bool(torch.equal(part.values[0, r0:r0 + n], V[r0:r0 + n]))
            assert bool(torch.equal(part.Q[0], vi.Q[0, r0:r0 + n]))
        else:
            assert bool(torch.allclose(part.values[0, r0:r0 + n], V[r0:r0 + n], rtol=2e-6, atol=0))
            assert bool(torch.allclose(part.Q[0], vi.Q[0, r0:r0 + n], rtol=2e-6, atol=0))


def test_edge_cases():
    import torch

    import colosseum_b200.dynamic_programming as dp
    import colosseum_b200.hardness as hd
    from colosseum_b200.batched_mdp import BatchedMDP
    from colosseum_b200.tables import MDPTables

    # single-state MDP: V* = r / (1 - gamma); diameter 0 targets itself
    T1 = np.ones((1, 2, 1), np.float32)
    R1 = np.array([[0.25, 0.5]], np.float32)
    Q, V = dp.discounted_value_iteration(T1, R1, 0.9, 1e-9, precision="f64")
    assert abs(V[0] - 0.5 / (1 - float(np.float32(0.9)))) < 1e-6
    assert hd.get_diameter(T1, False) == 0.0
    # empty batch of envs, and a batch that does not fill a warp
    tb = MDPTables.from_dense(np.full((3, 2, 3), 1 / 3, np.float32))
    for mode in ("dense_f32", "dense_f64"):
        e0 = BatchedMDP(tb, 0, mode=mode)
        e0.reset(); e0.step_async(None, auto_reset=True); torch.cuda.synchronize()
        e5 = BatchedMDP(tb, 5, mode=mode, seed=1)
        e5.reset()
        ts = e5.step(np.zeros(5, np.int32))
        assert ts.observation.shape == (5,) and int(e5.visits_s.sum()) == 10
    # zero-length horizon request and H = 1 episodic DP
    Tq = np.random.RandomState(0).dirichlet(np.ones(4), size=(4, 2)).astype(np.float32)
    Rq = np.random.RandomState(1).rand(4, 2).astype(np.float32)
    Qe, Ve = dp.episodic_value_iteration(1, Tq, Rq)
    assert Qe.shape == (2, 4, 2) and np.allclose(Qe[0], Rq) and (Ve[1] == 0).all()
