"""GPU probe: what the step kernel's time is made of (visitation counters, reward path, L2 state)"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from colosseum_b200.batched_mdp import BatchedMDP
from colosseum_b200.tables import MDPTables
g = np.load("tests/golden/inst_c2_deepsea30_prand.npz")
tb = MDPTables.from_golden(g)
N = 65536
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
gen = torch.Generator(device="cuda").manual_seed(1)
acts = [torch.randint(0, tb.A, (N,), dtype=torch.int32, device="cuda", generator=gen) for _ in range(8)]
def run(env, do_flush, n=200, random_actions=False):
    for i in range(20):
        env.step_async(None if random_actions else acts[i % 8], auto_reset=True)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for i, (a, b) in enumerate(ev):
        if do_flush: flush.zero_()
        a.record(); env.step_async(None if random_actions else acts[i % 8], auto_reset=True); b.record()
    torch.cuda.synchronize()
    return np.median([a.elapsed_time(b) for a, b in ev]) * 1e3
for name, kw in (("visits on", dict(track_visits=True)), ("visits off", dict(track_visits=False))):
    env = BatchedMDP(tb, N, mode="dense_f32", seed=1234, **kw); env.reset()
    print(f"{name:12s}: flushed {run(env, True):6.2f} us   warm {run(env, False):6.2f} us   warm+random actions {run(env, False, random_actions=True):6.2f} us")
# spread-out states (not DeepSea lockstep): uniform random states each step
env = BatchedMDP(tb, N, mode="dense_f32", seed=1, track_visits=True); env.reset()
env.state.copy_(torch.randint(0, tb.S, (N,), dtype=torch.int32, device="cuda"))
print(f"random states: flushed {run(env, True):6.2f} us   warm {run(env, False):6.2f} us")
# empty-ish kernel launch for reference: N=32 envs
e2 = BatchedMDP(tb, 32, mode="dense_f32", seed=1); e2.reset()
a32 = torch.zeros(32, dtype=torch.int32, device="cuda")
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(100)]
for a, b in ev:
    flush.zero_(); a.record(); e2.step_async(a32, auto_reset=True); b.record()
torch.cuda.synchronize()
print(f"32-env launch (event pair floor): {np.median([a.elapsed_time(b) for a, b in ev])*1e3:.2f} us")
