"""Multi-GPU check + timing of the row-sharded value iteration (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        scripts/multi_gpu_check.py [--S 16384] [--A 8] [--sweeps 50]

Checks that both transports ("nccl" all-gather, "fused" peer stores + barrier) reproduce the single-GPU sweep
exactly, then times them (device time, max over ranks)."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from colosseum_b200.dynamic_programming import BatchedValueIteration  # noqa: E402
from colosseum_b200.sharded import RowShardedValueIteration, shard_range  # noqa: E402
from colosseum_b200.synth import synth_dense_rows  # noqa: E402


def timed(fn, n):
    dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fn(n); b.record()
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b)], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()) / n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--S", type=int, default=8192)
    ap.add_argument("--A", type=int, default=8)
    ap.add_argument("--sweeps", type=int, default=50)
    ap.add_argument("--check-S", type=int, default=1000)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    out = {"world": world}
    # ---- correctness at a small, uneven-friendly size
    S, A = args.check_S, 3
    r0, r1 = shard_range(S, rank, world)
    T_rows, R_rows = synth_dense_rows(r0, r1 - r0, S, A, seed=5)
    T_full, R_full = synth_dense_rows(0, S, S, A, seed=5)
    assert torch.equal(T_full[r0:r1], T_rows) and torch.equal(R_full[r0:r1], R_rows)  # sharding-independent generator
    ref = BatchedValueIteration(T_full, R_full, gamma=0.99)
    ref.sweep(30)
    for transport in ("nccl", "fused"):
        try:
            vi = RowShardedValueIteration(T_rows, R_rows, S, gamma=0.99, transport=transport)
            vi.sweep(30)
            torch.cuda.synchronize()
            ok = bool(torch.equal(vi.values, ref.values.view(-1)))
            res = vi.residual()
            out[f"exact_{transport}"] = ok
            out[f"resid_{transport}"] = res
        except Exception as e:  # report, do not hide
            out[f"exact_{transport}"] = f"ERROR {type(e).__name__}: {e}"
    # ---- timing at the requested size
    S, A = args.S, args.A
    r0, r1 = shard_range(S, rank, world)
    T_rows, R_rows = synth_dense_rows(r0, r1 - r0, S, A, seed=7)
    bytes_per_sweep = 4.0 * S * A * S
    for transport in ("nccl", "fused"):
        try:
            vi = RowShardedValueIteration(T_rows, R_rows, S, gamma=0.99, transport=transport)
            vi.sweep(5)
            ms = timed(vi.sweep, args.sweeps)
            out[f"ms_per_sweep_{transport}"] = ms
            out[f"agg_GBps_{transport}"] = bytes_per_sweep / (ms * 1e-3) / 1e9
        except Exception as e:
            out[f"ms_per_sweep_{transport}"] = f"ERROR {type(e).__name__}: {e}"
    # local-only sweep (no exchange) for reference
    loc = BatchedValueIteration(T_rows[None], R_rows[None], gamma=0.99, row0=r0, S_total=S)
    loc.sweep(5)
    out["ms_per_sweep_no_exchange"] = timed(loc.sweep, args.sweeps)
    if rank == 0:
        print(json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
