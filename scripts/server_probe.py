"""Persistent step server: bit-identity with the launch-per-step path, lapse/restart, and end-to-end step time."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from colosseum_b200.batched_mdp import BatchedMDP, PipelinedBatchedMDP
import bench

tb = bench.load_c2_tables()
rng = np.random.default_rng(0)

# ---- correctness
for mode in ("dense_f32", "succ"):
    N = 5000
    acts = [torch.from_numpy(rng.integers(0, tb.A, N).astype(np.int32)).pin_memory() for _ in range(40)]
    ref = BatchedMDP(tb, N, mode=mode, seed=7, host_io=True); ref.reset()
    srv = BatchedMDP(tb, N, mode=mode, seed=7, host_io=True); srv.reset()
    buf = torch.zeros(N, dtype=torch.int32).pin_memory()
    srv.serve(buf, idle_timeout_ms=100)
    for i, a in enumerate(acts):
        o1, r1, s1 = [x.clone() for x in ref.step_host(a, auto_reset=True)]
        buf.copy_(a)
        o2, r2, s2 = srv.step_served()
        assert torch.equal(o1, o2) and torch.equal(s1, s2) and torch.equal(r1.view(torch.int32), r2.view(torch.int32)), (mode, i)
        if i == 20:
            time.sleep(0.4)  # let the server lapse: the next step restarts it
    srv.stop_serving()
    assert torch.equal(ref.visits_s, srv.visits_s) and torch.equal(ref.state, srv.state)
    print(mode, "served == launched for", len(acts), "steps (with one lapse/restart)", flush=True)

# ---- timing
def wall(f, n):
    for i in range(50): f()
    t0 = time.perf_counter()
    for i in range(n): f()
    return (time.perf_counter() - t0) / n * 1e6

for N in (32, 65536):
    buf = torch.from_numpy(rng.integers(0, tb.A, N).astype(np.int32)).pin_memory()
    env = BatchedMDP(tb, N, host_io=True); env.reset()
    t_launch = wall(lambda: env.step_host(buf, auto_reset=True), 500)
    env.serve(buf)
    t_srv = wall(env.step_served, 500)
    env.stop_serving()
    print(f"N={N:6d} step_host {t_launch:6.2f} us   served {t_srv:6.2f} us  ({N / t_srv / 1e3:.3f} G env-steps/s)", flush=True)

N = 65536
for groups in (2, 3, 4):
    env = PipelinedBatchedMDP(tb, N, groups=groups); env.reset()
    bufs = [torch.from_numpy(rng.integers(0, tb.A, env.sizes[g]).astype(np.int32)).pin_memory() for g in range(groups)]
    env.serve(bufs)
    for g in range(groups): env.send(g)
    K = 500
    t0 = time.perf_counter()
    for it in range(K):
        for g in range(groups):
            env.recv(g); env.send(g)
    dt = (time.perf_counter() - t0) / K * 1e6
    for g in range(groups): env.recv(g)
    env.stop_serving()
    print(f"pipelined served groups={groups}: {dt:6.2f} us/step ({N / dt / 1e3:.3f} G env-steps/s)", flush=True)
