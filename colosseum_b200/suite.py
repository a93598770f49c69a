"""The benchmark-suite workload (BASELINE.json configs[2], "C3"): many independent MDP instances of the reference's
families, each driven through the batched step (N parallel episodes x n steps) and then through the three hardness
measures -- exactly what `colosseum.hardness.analysis.compute_hardness_measure` + an agent/MDP loop do per instance
(colosseum/hardness/analysis.py:327-421, colosseum/mdp/base.py:996-1114), sharded over the GPUs of one box by
instance with no communication.

Instances travel in a SPARSE suite file (tests/golden/c3_suite.npz, written by tests/golden/make_c3_suite.py from the
unmodified reference): successor lists in the samplers' own order, reward distributions, start distribution, H, R.
The dense T is rebuilt here with the reference's own accumulation (float32 `T[s,a,s'] += p` in successor order,
colosseum/mdp/utils/mdp_creation.py:71-81) and is checked bit for bit against the CRC of the reference's `mdp.T`.
"""
import json
import time
import zlib
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

from .tables import MDPTables


@dataclass
class SuiteInstance:
    name: str
    tables: MDPTables
    R: np.ndarray  # f32 [S,A], the reference's mdp.R
    ref: Dict[str, float] = field(default_factory=dict)  # the reference's own answers (NaN = not recorded)
    nodes: Optional[np.ndarray] = None  # episodic: (h, s) pairs in the reference's episodic-graph order
    T_crc: int = 0

    @property
    def S(self):
        return self.tables.S

    @property
    def A(self):
        return self.tables.A

    @property
    def H(self):
        return self.tables.H

    @property
    def episodic(self):
        return self.tables.H > 0

    def check_T(self):
        """the rebuilt dense T is bit-identical to the reference's mdp.T"""
        return zlib.crc32(np.ascontiguousarray(self.tables.T, np.float32).tobytes()) == self.T_crc


def load_suite(path, only=None) -> List[SuiteInstance]:
    z = np.load(path, allow_pickle=False)
    names = json.loads(str(z["names"]))
    out = []
    for i, name in enumerate(names):
        if only is not None and name not in only and i not in only:
            continue
        k = f"i{i}_"
        kinds = [(a, tuple(b)) for a, b in json.loads(str(z[k + "rew_kinds"]))]
        tb = MDPTables.from_successors(int(z[k + "S"]), int(z[k + "A"]), z[k + "succ_idx"].astype(np.int32),
                                       z[k + "succ_prob"], z[k + "succ_len"].astype(np.int32),
                                       z[k + "rew_cls"].astype(np.int32), kinds, z[k + "start_idx"], z[k + "start_prob"],
                                       H=int(z[k + "H"]), rewards_range=tuple(z[k + "rewards_range"]))
        ref = {m: float(z[k + m]) for m in ("value_norm", "gaps", "diameter", "cached_diameter", "cached_value_norm")}
        nodes = None
        if (k + "reach_h") in z.files:
            nodes = np.stack([z[k + "reach_h"].astype(np.int64), z[k + "reach_s"].astype(np.int64)], 1)
        out.append(SuiteInstance(name, tb, np.asarray(z[k + "R"], np.float32), ref, nodes, int(z[k + "T_crc"])))
    return out


def suite_files(golden_dir):
    import os

    files = [os.path.join(golden_dir, "c3_suite.npz")]
    more = os.path.join(golden_dir, "c3_suite_seeds.npz")
    if os.path.isfile(more):
        files.append(more)
    return files


def suite_size(golden_dir) -> int:
    return sum(len(json.loads(str(np.load(f, allow_pickle=False)["names"]))) for f in suite_files(golden_dir))


def load_suite_all(golden_dir, indices=None) -> List[SuiteInstance]:
    """SURVEY section 8d's C3 list: every gin parameter set of the six benchmark folders x seeds 0..10, seed-major --
    c3_suite.npz (the 80 sets of the four benchmarks at seed 0, with the reference's own hardness answers) followed
    by c3_suite_seeds.npz (the 14 quick-test sets at seed 0, then all 94 sets at seeds 1..10) when that file exists.
    `indices`: positions in that list to load (each instance rebuilds its dense T on the host: load what a rank needs);
    the result follows the order of `indices`."""
    sizes, files = [], suite_files(golden_dir)
    for f in files:
        sizes.append(len(json.loads(str(np.load(f, allow_pickle=False)["names"]))))
    if indices is None:
        out = []
        for f in files:
            out += load_suite(f)
        return out
    want = {}
    for i in indices:
        j, k = int(i), 0
        while j >= sizes[k]:
            j -= sizes[k]
            k += 1
        want.setdefault(k, set()).add(j)
    loaded = {}
    for k, js in want.items():
        names = json.loads(str(np.load(files[k], allow_pickle=False)["names"]))
        for inst, j in zip(load_suite(files[k], only=js), sorted(js)):
            assert inst.name == names[j]
            loaded[(k, j)] = inst
    out = []
    for i in indices:
        j, k = int(i), 0
        while j >= sizes[k]:
            j -= sizes[k]
            k += 1
        out.append(loaded[(k, j)])
    return out


def hardness_of_instance(inst: SuiteInstance, precision="f64", max_cf_bytes=8 << 30, diameter=True, timings=None):
    """diameter, environmental value norm and sum of reciprocal sub-optimality gaps of one instance on the GPU, with
    the property layer's choices of the reference: continuous MDPs use T, R and discounted VI
    (colosseum/mdp/base.py:635-647, :1042-1100); episodic ones use the episodic tensor for the diameter
    (base.py:996-1016), backward induction + reachable (h,s) pairs for the gaps (base.py:1018-1040) and the
    continuous form for the value norm (base.py:1049-1056, base_finite.py:167-178).  Returns a dict."""
    import torch

    from . import dynamic_programming as dp
    from . import episodic_forms as ef
    from . import hardness as hd

    def lap(name, _t=[None]):
        if timings is not None:
            torch.cuda.synchronize()
            now = time.perf_counter()
            if _t[0] is not None:
                timings[name] = timings.get(name, 0.0) + now - _t[0]
            _t[0] = now

    lap(None)
    tb = inst.tables
    T = torch.from_numpy(tb.T).cuda()
    R = torch.from_numpy(inst.R).cuda()
    start_p = np.diff(tb.start_cum, prepend=0.0)
    out = {}
    lap("upload")
    eps = 1e-9 if precision == "f64" else 1e-5
    deterministic = bool((tb.succ_len == 1).all()) and all(k == "deterministic" for k, _ in tb.rew_kinds)
    if not inst.episodic:
        Q, V = dp.discounted_value_iteration(T, R, 0.99, eps, precision=precision)
        lap("vi")
        out["gaps"] = hd.get_sum_reciprocals_suboptimality_gaps(Q, V)
        lap("gaps")
        out["value_norm"] = 0.0 if deterministic else hd.calculate_norm_discounted(T, V, precision=precision)
        lap("value_norm")
        if diameter:
            out["diameter"], out["diameter_sweeps"] = hd.get_diameter(T, False, precision=precision, return_sweeps=True)
            lap("diameter")
    else:
        H = tb.H
        Q, V = dp.episodic_value_iteration(H, T, R, precision=precision)
        lap("vi")
        T_epi, _, reach = ef.get_episodic_transition_matrix_and_rewards(H, T, R, tb.start_idx, start_p, return_reach=True)
        nodes = inst.nodes if inst.nodes is not None else torch.nonzero(reach).cpu().numpy()
        lap("T_epi")
        out["gaps"] = hd.get_sum_reciprocals_suboptimality_gaps(Q, V, np.asarray(nodes))
        lap("gaps")
        n = len(nodes)
        if deterministic:
            out["value_norm"] = 0.0
        elif 4 * n * n * tb.A <= max_cf_bytes:
            T_cf, R_cf = ef.get_continuous_form_episodic_transition_matrix_and_rewards(H, T, R, tb.start_idx, start_p,
                                                                                       nodes=nodes)
            lap("T_cf")
            # V* of the continuous form from its structure (a scalar fixed point + backward inductions over T)
            # instead of ~2,000 sweeps over the n x A x n tensor; same fixed point (tests/test_gpu_hardness.py)
            V_cf = ef.continuous_form_optimal_values(H, T, R, tb.start_idx, start_p, nodes=nodes, gamma=0.99,
                                                     epsilon=eps, precision=precision)
            lap("vi_cf")
            out["value_norm"] = hd.calculate_norm_discounted(T_cf, V_cf, precision=precision)
            del T_cf, R_cf
            lap("value_norm")
        else:  # the reference raises "Its continuous form is too large" (mdp_creation.py:152-155)
            out["value_norm"] = float("nan")
        if diameter:
            out["diameter"], out["diameter_sweeps"] = hd.get_diameter(T_epi, True, precision=precision, return_sweeps=True)
            lap("diameter")
    return out


def run_instance(inst: SuiteInstance, n_envs=1024, n_steps=1000, seed=0, mode="succ", precision="f64", diameter=True):
    """one C3 work item: `n_envs` parallel episodes x `n_steps` random-agent steps (BaseMDP.random_steps with
    auto_reset, base.py:1319-1355), then the hardness measures.  Returns (results dict, seconds per phase)."""
    import torch

    from .batched_mdp import BatchedMDP

    t0 = time.perf_counter()
    env = BatchedMDP(inst.tables, n_envs, mode=mode, seed=seed)
    env.reset()
    env.random_steps_fused(n_steps, auto_reset=True)  # one launch: every env walks n_steps with random actions
    visits = env.visits_s
    total = int(visits.sum().item())  # syncs
    t1 = time.perf_counter()
    res = hardness_of_instance(inst, precision=precision, diameter=diameter)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    res["visits_total"] = total
    res["mean_reward_last_step"] = float(torch.nan_to_num(env.reward, nan=0.0).mean().item())
    return res, {"step_s": t1 - t0, "hardness_s": t2 - t1}


def run_many(work, n_workers=4, **kw):
    """Run C3 work items `work` = [(SuiteInstance, seed), ...] on the current device with `n_workers` host threads, each
    on its own CUDA stream.  The instances are independent (no communication), and a single one
    keeps only a few SMs busy -- the on-chip solvers give one CTA (or one CTA per few targets) to an MDP with a few
    hundred states -- so several instances in flight fill the GPU; the C ABI is stream-ordered and releases the GIL
    while it waits.  Returns the list of (results, timings) in the order of `work`."""
    import threading
    from concurrent.futures import ThreadPoolExecutor

    import torch

    dev = torch.cuda.current_device()
    local = threading.local()

    def one(item):
        inst, seed = item
        if not hasattr(local, "stream"):
            torch.cuda.set_device(dev)
            local.stream = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(local.stream):
            out = run_instance(inst, seed=seed, **kw)
            local.stream.synchronize()
        return out

    if n_workers <= 1:
        return [one(w) for w in work]
    # the workers alternate between short stretches of Python and GIL-releasing C calls; with CPython's default 5 ms
    # switch interval a thread coming back from a C call waits up to 5 ms for the GIL while another one runs Python
    # (measured: 4 workers were no faster than 1).  A short interval hands the GIL over at once.
    import os
    import sys

    old = sys.getswitchinterval()
    sys.setswitchinterval(float(os.environ.get("COLO_SWITCH_INTERVAL", "5e-5")))
    try:
        with ThreadPoolExecutor(max_workers=n_workers) as pool:
            return list(pool.map(one, work))
    finally:
        sys.setswitchinterval(old)


def _cost_ms(episodic: bool, n: int) -> float:
    # least-squares fit of the measured wall time of an instance (step + hardness phases, 8 instances in flight on one
    # B200, profiles/r2_c3_phases.txt): a fixed ~10-12 ms of launch / synchronise round trips plus a mild size term
    return 12.5 + 0.0029 * n if episodic else 10.4 + 0.0133 * n


def instance_cost(inst: "SuiteInstance") -> float:
    """estimated milliseconds of the C3 work item of an instance: grows with the reachable (h, s) nodes of an episodic
    MDP (its episodic / continuous-form tensors, mdp/base.py:996-1056), with the states of a continuous one"""
    ep = inst.episodic and inst.nodes is not None
    return _cost_ms(ep, len(inst.nodes) if ep else inst.S)


def suite_costs(golden_dir) -> np.ndarray:
    """`instance_cost` of every instance of the C3 list from the fixtures' metadata alone (S, number of reachable (h, s)
    nodes): cheap enough for every rank to compute for the whole list"""
    out = []
    for f in suite_files(golden_dir):
        z = np.load(f, allow_pickle=False)
        for i in range(len(json.loads(str(z["names"])))):
            k = f"i{i}_"
            ep = (k + "reach_h") in z.files
            out.append(_cost_ms(ep, int(z[k + "reach_h"].shape[0]) if ep else int(z[k + "S"])))
    return np.asarray(out)


def shard_instances(n_items: int, rank: int, world: int, costs=None):
    """Indices of the C3 instance list (cycled to n_items) a rank owns; no communication.  With `costs` (one per list
    entry, `suite_costs`): the items sorted by decreasing cost are dealt to the ranks in snake order (0..w-1, w-1..0, ...),
    so every rank gets the same number of instances (+-1), nearly the same total cost, and its share already longest
    first -- a contiguous range gives every rank a different partial cycle of the 94 parameter sets.  Without costs:
    i % world == rank."""
    n_items, rank, world = int(n_items), int(rank), int(world)
    if costs is None:
        return list(range(rank, n_items, world))
    c = np.asarray(costs, np.float64)
    order = sorted(range(n_items), key=lambda i: (-c[i % len(c)], i))
    mine = []
    for j, i in enumerate(order):
        r = j % world
        if (j // world) % 2:
            r = world - 1 - r
        if r == rank:
            mine.append(i)
    return mine


def longest_first(work):
    """`work` = [(SuiteInstance, seed), ...] ordered by decreasing `instance_cost`: the workers of `run_many_native` pull
    from one queue, and the most expensive instance is a large part of a 128-instance shard's wall time"""
    return sorted(work, key=lambda w: -instance_cost(w[0]))


def release_caches():
    """frees the large device buffers colo_suite_run's workers keep between instances and calls"""
    from . import _cabi

    _cabi.check(_cabi.lib().colo_suite_release_caches(), "colo_suite_release_caches")


def run_many_native(work, n_workers=8, n_envs=1024, n_steps=1000, diameter=True, max_cf_bytes=8 << 30, eps=1e-9):
    """The same C3 work items as `run_many`, scheduled by the library (colo_suite_run, csrc/suite_runner.cu): `n_workers`
    C++ host threads with one CUDA stream each run whole instances -- every step the public C-ABI entry point the Python
    path calls, fp64 accumulation, iterations to `eps` -- so many small solves are in flight at once with no
    interpreter between them.  `work` = [(SuiteInstance, seed), ...]: one seed per call (the first item's).  Returns
    the list of (results, timings) in the order of `work`, as `run_many` does."""
    import ctypes as C

    from . import _cabi

    _cabi.require_cuda()
    n = len(work)
    arr = (_cabi.SuiteInstance * max(n, 1))()
    keep = []

    def p(a, dt):
        a = np.ascontiguousarray(a, dt)
        keep.append(a)
        return a.ctypes.data

    for i, (inst, _) in enumerate(work):
        tb = inst.tables
        c = arr[i]
        c.S, c.A, c.H, c.K = tb.S, tb.A, tb.H, tb.succ_cum.shape[-1]
        c.T, c.R = p(tb.T, np.float32), p(inst.R, np.float32)
        c.succ_cum, c.succ_idx = p(tb.succ_cum, np.float64), p(tb.succ_idx, np.int32)
        c.succ_len, c.rew_cls_succ = p(tb.succ_len, np.int32), p(tb.rew_cls_succ, np.int32)
        c.rew_q = p(tb.rew_q, np.float32)
        c.n_cls, c.nq = tb.rew_q.shape
        c.rmin, c.rmax = tb.rmin, tb.rmax
        c.start_idx, c.start_cum = p(tb.start_idx, np.int32), p(tb.start_cum, np.float64)
        c.start_prob = p(np.diff(tb.start_cum, prepend=0.0), np.float64)
        c.n_start = tb.n_start
        if inst.episodic:
            assert inst.nodes is not None, "episodic instances carry the reference's (h, s) node order"
            c.node_h, c.node_s = p(inst.nodes[:, 0], np.int32), p(inst.nodes[:, 1], np.int32)
            c.n_nodes = len(inst.nodes)
        c.deterministic = int(bool((tb.succ_len == 1).all()) and all(k == "deterministic" for k, _ in tb.rew_kinds))
    cfg = _cabi.SuiteConfig(int(n_envs), int(n_steps), int(work[0][1]) if n else 0, float(eps), int(max_cf_bytes),
                            int(bool(diameter)))
    res = (_cabi.SuiteResult * max(n, 1))()
    rc = _cabi.lib().colo_suite_run(arr, n, C.byref(cfg), res, int(n_workers))
    if rc != 0:
        bad = [(work[i][0].name, res[i].status, res[i].error.decode(errors="replace")) for i in range(n) if res[i].status]
        raise _cabi.ColosseumB200Error(f"colo_suite_run: rc={rc}: {bad[:3]}")
    out = []
    for i in range(n):
        r = res[i]
        out.append(({"gaps": r.gaps, "value_norm": r.value_norm, "diameter": r.diameter,
                     "diameter_sweeps": int(r.diameter_sweeps) if r.diameter_sweeps == r.diameter_sweeps else 0,
                     "visits_total": int(r.visits_total), "mean_reward_last_step": r.mean_reward_last_step},
                    {"step_s": r.step_s, "hardness_s": r.hardness_s}))
    return out
