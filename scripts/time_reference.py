"""Times the UNMODIFIED Python reference (/root/reference, imported under oracle/ref_shim) in the BUILD container,
once, and writes profiles/reference_python_timing.json (C1 random-agent interaction +
episodic VI, C2 mdp.step).  /root/reference does not exist on the GPU box, so bench.py only QUOTES this file
(`cpu_baseline.sample`); nothing here runs at bench time.

    python scripts/time_reference.py
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.reference_import import import_reference  # noqa: E402

colosseum = import_reference()
from colosseum.dynamic_programming import episodic_value_iteration  # noqa: E402
from colosseum.mdp.deep_sea import DeepSeaContinuous  # noqa: E402
from colosseum.mdp.river_swim import RiverSwimEpisodic  # noqa: E402

out = {"cores_used": 1, "cpu": open("/proc/cpuinfo").read().split("model name")[1].split(":")[1].split("\n")[0].strip(),
       "host_cores": os.cpu_count(), "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()),
       "how": "single Python process, time.perf_counter, best of 3"}

# ---- C1: RiverSwimEpisodic size 5, the quick-test gin parameters
c1 = RiverSwimEpisodic(seed=0, size=5, p_lazy=0.1, make_reward_stochastic=True, randomize_actions=False,
                       sub_optimal_distribution=("beta", (2.4, 24.0)), optimal_distribution=("beta", (0.01, 0.11)),
                       other_distribution=("beta", (2.4, 249.0)))
c1.reset()
c1.random_steps(2000, auto_reset=True)
best = 1e9
for _ in range(3):
    t0 = time.perf_counter()
    c1.random_steps(10_000, auto_reset=True)
    best = min(best, time.perf_counter() - t0)
out["c1_random_steps_per_s"] = 10_000 / best
T, R = c1.T, c1.R
episodic_value_iteration(c1.H, T, R)
best = 1e9
for _ in range(3):
    t0 = time.perf_counter()
    for _ in range(200):
        Q, V = episodic_value_iteration(c1.H, T, R)
    best = min(best, (time.perf_counter() - t0) / 200)
out["c1_episodic_vi_us"] = best * 1e6
out["c1_V0"] = [float(x) for x in V[0]]

# ---- C2: DeepSeaContinuous size 30, p_rand = 0.1, mdp.step with supplied random actions
c2 = DeepSeaContinuous(seed=0, size=30, p_rand=0.1)
c2.reset()
rs = np.random.RandomState(0)
acts = rs.randint(0, c2.n_actions, size=20_000)
for a in acts[:2000]:
    c2.step(int(a))
best = 1e9
for _ in range(3):
    t0 = time.perf_counter()
    for a in acts:
        c2.step(int(a))
    best = min(best, time.perf_counter() - t0)
out["c2_step_per_s"] = len(acts) / best
out["c2_S_A"] = [int(c2.n_states), int(c2.n_actions)]

# ---- C4 shape: the reference's dense numba kernel (infinite_horizon.py:121-142), one MDP S=512 A=4, one core
from colosseum.dynamic_programming.infinite_horizon import discounted_value_iteration  # noqa: E402

rs = np.random.RandomState(0)
T4 = rs.dirichlet(np.ones(512) * 0.05, size=(512, 4)).astype(np.float32)
R4 = rs.uniform(0, 1, size=(512, 4)).astype(np.float32)
discounted_value_iteration(T4, R4, 0.99, 1e-3)  # JIT
best = 1e9
for _ in range(5):
    t0 = time.perf_counter()
    Q4, V4 = discounted_value_iteration(T4, R4, 0.99, 1e-3)
    best = min(best, time.perf_counter() - t0)
# count the in-place sweeps the reference needed with the oracle's restatement of the same iterate
from oracle import oracle as orc  # noqa: E402

_, _, n_sweeps = orc.discounted_gs_f32(T4, R4, gamma=0.99, eps=1e-3)
out["c4_numba_solve_ms"] = best * 1e3
out["c4_numba_sweeps"] = int(n_sweeps)
out["c4_numba_mdp_sweeps_per_s"] = int(n_sweeps) / best

p = os.path.join(ROOT, "profiles", "reference_python_timing.json")
json.dump(out, open(p, "w"), indent=1)
print(json.dumps(out))
