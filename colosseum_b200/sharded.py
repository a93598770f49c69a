"""Multi-GPU sharding of the hot path on one 8xB200 box, one process per GPU.

  * parallel envs        -- `shard_range` + BatchedMDP(env_offset=...): no communication.
  * independent MDPs     -- `shard_range` over the instance axis: no communication.
  * one large MDP        -- `RowShardedValueIteration`: rank r owns rows [r*S/g, (r+1)*S/g) of T; every sweep each
                            rank backs up its rows from the full V and the new V rows are all-gathered.  Two
                            transports: "nccl" (torch.distributed all_gather_into_tensor over NVLink/NVSwitch) and
                            "fused" (the backup kernel itself stores its V rows into every peer's V buffer through
                            torch symmetric memory; one signal-pad barrier per sweep replaces the collective).
The shard arithmetic is pure Python and is covered by world_size-2 gloo tests on CPU.
"""
import numpy as np


def shard_range(n_items: int, rank: int, world: int):
    """contiguous, balanced split: the first n_items % world ranks get one extra item"""
    base, extra = divmod(int(n_items), int(world))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_sizes(n_items: int, world: int):
    return [shard_range(n_items, r, world)[1] - shard_range(n_items, r, world)[0] for r in range(world)]


def allgather_rows_host(local_rows: np.ndarray, n_items: int, group=None):
    """gloo/CPU reference of the V all-gather (uneven shards allowed): used by the CPU tests of the host logic"""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    sizes = shard_sizes(n_items, world)
    mx = max(sizes)
    local = torch.from_numpy(np.ascontiguousarray(local_rows))
    padded = torch.zeros(mx, dtype=local.dtype)
    padded[: local.numel()] = local
    bufs = [torch.zeros(mx, dtype=local.dtype) for _ in sizes]  # equal-size exchange; the padding is dropped
    dist.all_gather(bufs, padded, group=group)
    return torch.cat([b[:s] for b, s in zip(bufs, sizes)]).numpy()


class RowShardedValueIteration:
    """Value iteration on ONE dense MDP whose T is row-sharded over the ranks of `group` (config C5)."""

    def __init__(self, T_rows, R_rows, S, gamma=0.99, transport="nccl", group=None, precision="f32", local_vi=None):
        """local_vi: the object that backs up this rank's rows -- anything with `.sweep(1)`, `.values` ([1, S] tensor
        whose rows [row0, row1) the sweep rewrites from the full vector), `.residual()` and `.cur`; default: the CUDA
        BatchedValueIteration on (T_rows, R_rows).  The exchange logic below only sees that interface, which is how the
        world-size-2 gloo tests drive THIS class on CPU tensors (tests/test_sharding_cpu.py)."""
        import torch
        import torch.distributed as dist

        self.dist, self.torch = dist, torch
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.S = int(S)
        self.row0, self.row1 = shard_range(S, self.rank, self.world)
        assert T_rows is None or T_rows.shape[0] == self.row1 - self.row0, "T_rows must hold exactly this rank's rows"
        self.transport = transport
        if local_vi is None:
            from .dynamic_programming import BatchedValueIteration

            local_vi = BatchedValueIteration(T_rows[None], R_rows[None], gamma=gamma, precision=precision,
                                             row0=self.row0, S_total=S)
        self.vi = local_vi
        self.sizes = shard_sizes(S, self.world)
        self.even = len(set(self.sizes)) == 1
        if transport == "fused":
            self._setup_symmetric()

    def _setup_symmetric(self):
        """V ping/pong buffers in torch symmetric memory; every rank learns every peer's pointers."""
        import torch.distributed._symmetric_memory as symm

        torch = self.torch
        vd = self.vi.V[0].dtype
        self._symm = []
        peers = []
        for i in range(2):
            buf = symm.empty((1, self.S), dtype=vd, device="cuda")
            hdl = symm.rendezvous(buf, self.group if self.group is not None else self.dist.group.WORLD)
            buf.zero_()
            self.vi.V[i] = buf
            ptrs = [int(hdl.buffer_ptrs[r]) for r in range(self.world) if r != self.rank]
            peers.append(torch.tensor(ptrs, dtype=torch.int64, device="cuda"))
            self._symm.append(hdl)
        self.vi.set_peers(peers)
        self._symm[0].barrier(channel=0)

    def sweep(self, n=1):
        torch, dist = self.torch, self.dist
        for _ in range(n):
            self.vi.sweep(1)
            V = self.vi.values  # [1,S]: rows [row0,row1) are new; the rest must come from the peers
            if self.transport == "nccl":
                if self.even:
                    dist.all_gather_into_tensor(V.view(-1), V.view(-1)[self.row0:self.row1].clone(), group=self.group)
                else:  # uneven shards: equal-size padded exchange, then unpack
                    mx = max(self.sizes)
                    pad = torch.zeros(mx, dtype=V.dtype, device=V.device)
                    pad[: self.row1 - self.row0] = V.view(-1)[self.row0:self.row1]
                    out = torch.empty(self.world * mx, dtype=V.dtype, device=V.device)
                    dist.all_gather_into_tensor(out, pad, group=self.group)
                    for r in range(self.world):
                        a, b = shard_range(self.S, r, self.world)
                        V.view(-1)[a:b] = out[r * mx: r * mx + (b - a)]
            else:
                # peers' rows were written by their kernels; one barrier makes them visible before the next sweep
                self._symm[self.vi.cur].barrier(channel=0)

    @property
    def values(self):
        return self.vi.values.view(-1)

    def residual(self):
        """global max|dV| of the sweeps since the last call (max over ranks of the local-row residuals)"""
        r = self.vi.residual()
        self.dist.all_reduce(r, op=self.dist.ReduceOp.MAX, group=self.group)
        return float(r.item())
