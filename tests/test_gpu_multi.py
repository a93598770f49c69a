"""GPU, two devices (skipped on a single-GPU box): the row-sharded value iteration of config C5
(colosseum_b200.sharded.RowShardedValueIteration: NCCL all-gather of V, and the fused variant whose backup kernel
stores its V rows into the peer's buffer) against the single-GPU solver on the same MDP -- bit for bit while a shard
keeps the warp-per-state mapping, to rounding otherwise -- and env sharding with global Philox counters."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, S, A, transport, q):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from colosseum_b200.batched_mdp import BatchedMDP
        from colosseum_b200.dynamic_programming import BatchedValueIteration
        from colosseum_b200.sharded import RowShardedValueIteration, shard_range
        from colosseum_b200.synth import synth_dense_rows
        from colosseum_b200.tables import MDPTables

        r0, r1 = shard_range(S, rank, world)
        T_rows, R_rows = synth_dense_rows(r0, r1 - r0, S, A, seed=7)
        sh = RowShardedValueIteration(T_rows, R_rows, S, gamma=0.99, transport=transport)
        sh.sweep(12)
        T_all, R_all = synth_dense_rows(0, S, S, A, seed=7)
        one = BatchedValueIteration(T_all, R_all, gamma=0.99, precision="f32")
        one.sweep(12)
        got, ref = sh.values.view(-1), one.values.view(-1)
        err = float(((got - ref).abs() / ref.abs().clamp_min(1e-30)).max())
        # env sharding: each rank steps its slice with env_offset = global index; the union equals the unsharded batch
        rs = np.random.RandomState(0)
        Tm = rs.dirichlet(np.ones(48) * 0.2, size=(48, 3)).astype(np.float32)
        tb = MDPTables.from_dense(Tm, start_idx=np.arange(48), start_prob=np.ones(48) / 48)
        N = 1000
        e0, e1 = shard_range(N, rank, world)
        env = BatchedMDP(tb, e1 - e0, mode="dense_f32", seed=9, env_offset=e0)
        env.reset()
        for _ in range(5):
            env.step_async(None, auto_reset=True)
        full = BatchedMDP(tb, N, mode="dense_f32", seed=9)
        full.reset()
        for _ in range(5):
            full.step_async(None, auto_reset=True)
        ok_env = bool((env.state == full.state[e0:e1]).all())
        q.put((rank, err, ok_env))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("transport", ["nccl", "fused"])
@pytest.mark.parametrize("S", [4096, 1001])  # even and uneven shards
def test_row_sharded_vi_two_gpus(transport, S):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, S, 4, transport, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, err, ok_env in res:
        assert err <= 2e-6, (rank, err)
        assert ok_env, rank
