"""CPU (gloo, world_size 2): the host-side logic of the multi-GPU paths -- shard arithmetic, the V all-gather of a
row-sharded sweep, env/instance sharding with disjoint Philox streams.  The per-rank arithmetic is done by the
oracle here (no GPU); what is under test is the partition / exchange logic of colosseum_b200.sharded."""
import os
import socket

import numpy as np
import pytest

from colosseum_b200.sharded import allgather_rows_host, shard_range, shard_sizes


def test_shard_range_partitions():
    for n in (0, 1, 7, 8, 9, 40000, 65536, 1024):
        for w in (1, 2, 3, 4, 8):
            rngs = [shard_range(n, r, w) for r in range(w)]
            assert rngs[0][0] == 0 and rngs[-1][1] == n
            assert all(rngs[i][1] == rngs[i + 1][0] for i in range(w - 1))
            sizes = shard_sizes(n, w)
            assert sum(sizes) == n and max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _HostRowsVI:
    """the local-VI interface of RowShardedValueIteration on CPU tensors: this rank's rows backed up from the full V by
    the oracle's fp64 arithmetic (test double of the CUDA BatchedValueIteration; the class under test is the sharded
    solver's partition / exchange logic, which only sees this interface)"""

    def __init__(self, T_rows, R_rows, S, row0, gamma):
        import torch

        self.T, self.R, self.row0, self.gamma = T_rows.astype(np.float64), R_rows.astype(np.float64), row0, gamma
        self.V = [torch.zeros((1, S), dtype=torch.float64) for _ in range(2)]
        self.cur = 0
        self._res = 0.0

    @property
    def values(self):
        return self.V[self.cur]

    def sweep(self, n=1):
        import torch

        for _ in range(n):
            v = self.V[self.cur].view(-1).numpy()
            new = (self.R + self.gamma * (self.T @ v)).max(-1)
            nxt = 1 - self.cur
            out = self.V[nxt].view(-1)
            r1 = self.row0 + len(new)
            self._res = max(self._res, float(np.abs(new - v[self.row0:r1]).max()))
            out[self.row0:r1] = torch.from_numpy(new)
            self.cur = nxt

    def residual(self):
        import torch

        r, self._res = self._res, 0.0
        return torch.tensor([r], dtype=torch.float64)


def _worker(rank, world, port, S, A, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as orc
        from colosseum_b200.sharded import RowShardedValueIteration

        # ---- the REAL RowShardedValueIteration (transport "nccl" == torch.distributed collectives, gloo here), even
        # and uneven shards, driven through its public interface with a host-side local VI
        rs0 = np.random.RandomState(1)
        Tm = rs0.dirichlet(np.ones(S) * 0.1, size=(S, A)).astype(np.float32)
        Rm = rs0.uniform(0, 1, (S, A)).astype(np.float32)
        a0, a1 = shard_range(S, rank, world)
        sh = RowShardedValueIteration(None, None, S, gamma=0.9, transport="nccl",
                                      local_vi=_HostRowsVI(Tm[a0:a1], Rm[a0:a1], S, a0, 0.9))
        assert (sh.row0, sh.row1) == (a0, a1)
        sh.sweep(25)
        _, V_ref25 = orc.jacobi_sweeps_f64(Tm, Rm, np.zeros(S), 25, gamma=0.9)
        ok_class = bool(np.array_equal(sh.values.numpy(), V_ref25))
        res = sh.residual()  # max over ranks of the local residuals since the start
        _, V_ref1 = orc.jacobi_sweeps_f64(Tm, Rm, np.zeros(S), 1, gamma=0.9)
        ok_class = ok_class and abs(res - float(np.abs(V_ref1).max())) < 1e-12

        rs = np.random.RandomState(0)  # every rank builds the same MDP, keeps only its rows
        T = rs.dirichlet(np.ones(S) * 0.1, size=(S, A)).astype(np.float32)
        R = rs.uniform(0, 1, (S, A)).astype(np.float32)
        r0, r1 = shard_range(S, rank, world)
        V = np.zeros(S)
        for _ in range(25):  # row-sharded synchronous sweeps: local rows from the full V, then all-gather
            Q_rows = R[r0:r1].astype(np.float64) + 0.9 * (T[r0:r1].astype(np.float64) @ V)
            V = allgather_rows_host(Q_rows.max(-1), S)
        _, V_ref = orc.jacobi_sweeps_f64(T, R, np.zeros(S), 25, gamma=0.9)
        ok_vi = bool(np.array_equal(V, V_ref))
        # env sharding: each rank steps its slice with env0 = global offset; concatenation == unsharded run
        from colosseum_b200.tables import MDPTables

        tb = MDPTables.from_dense(T, start_idx=np.arange(S), start_prob=np.ones(S) / S)
        cdf = orc.build_dense_cdf(T, ld=tb.ld)
        ht = orc.HostTables(S, A, cdf=cdf, rew_q=tb.rew_q, start_cum=tb.start_cum, start_idx=tb.start_idx)
        N = 1001
        e0, e1 = shard_range(N, rank, world)
        st_l, h_l, ty_l, _ = orc.env_reset(ht, e1 - e0, seed=9, t=0, env0=e0)
        for t in range(1, 6):
            orc.env_step(ht, 0, st_l, h_l, ty_l, action=None, seed=9, t=t, env0=e0)
        full = allgather_rows_host(st_l, N)
        st_f, h_f, ty_f, _ = orc.env_reset(ht, N, seed=9, t=0)
        for t in range(1, 6):
            orc.env_step(ht, 0, st_f, h_f, ty_f, action=None, seed=9, t=t)
        q.put((rank, ok_vi and ok_class, bool(np.array_equal(full, st_f))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("S", [64, 37])  # even and uneven shards
def test_row_sharded_vi_and_env_sharding_gloo(S):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, S, 3, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok_vi and ok_env for _, ok_vi, ok_env in res), res
