"""The C3 benchmark suite (BASELINE.json configs[2]): the reference's benchmark MDP instances in sparse form.

CPU part: the fixture loads, and the dense T rebuilt from the successor lists is bit-identical to the reference's
`mdp.T` (CRC recorded by tests/golden/make_c3_suite.py).  GPU part: step + hardness measures of every instance
against the values the unmodified reference produced (and its cached_hardness_measures files)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from colosseum_b200.suite import load_suite
from oracle import oracle as orc

SUITE = os.path.join(GOLDEN, "c3_suite.npz")


def rel(a, b):
    return abs(a - b) / max(abs(b), 1e-12)


@pytest.fixture(scope="module")
def suite():
    return load_suite(SUITE)


def test_suite_fixture_is_the_reference_benchmark(suite):
    assert len(suite) == 80
    fams = {i.name.split(".")[0] for i in suite}
    assert len(fams) == 14  # seven families x {Continuous, Episodic}
    for inst in suite:
        assert inst.check_T(), inst.name  # rebuilt T == reference mdp.T, bit for bit
        T = inst.tables.T
        assert np.allclose(T.sum(-1), 1.0, atol=1e-5)  # mdp_creation.py:93
        # R recomputed from the tables (sum_s' p E[r], mdp_creation.py:71-81) matches the reference's mdp.R
        np.testing.assert_allclose(inst.tables.expected_rewards(), inst.R, rtol=2e-6, atol=1e-7)
        assert (inst.nodes is not None) == inst.episodic


def test_shard_scheduling(suite):
    """C3 sharding: every index exactly once, equal counts, nearly equal cost per rank, each rank's share longest first"""
    from colosseum_b200.suite import instance_cost, longest_first, shard_instances, suite_costs

    costs = suite_costs(GOLDEN)
    assert len(costs) >= 80 and all(costs[i] == instance_cost(inst) for i, inst in enumerate(suite))
    for n, world in ((1024, 8), (len(costs), 8), (128, 1), (256, 2), (7, 3), (2, 4)):
        for cs in (None, costs):
            parts = [shard_instances(n, r, world, costs=cs) for r in range(world)]
            assert sorted(i for p in parts for i in p) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
        tot = [sum(costs[i % len(costs)] for i in p) for p in parts]
        if n >= 128:
            assert max(tot) <= 1.01 * min(tot), (n, world, tot)
        for p in parts:
            c = [costs[i % len(costs)] for i in p]
            assert c == sorted(c, reverse=True)
    work = longest_first([(inst, 0) for inst in suite])
    c = [instance_cost(w[0]) for w in work]
    assert c == sorted(c, reverse=True) and sorted(id(w[0]) for w in work) == sorted(id(i) for i in suite)


def test_oracle_on_small_suite_instances(suite):
    """the CPU oracle against the reference's recorded measures on the small instances (keeps the CPU suite fast)"""
    n = 0
    for inst in suite:
        if inst.episodic or inst.S > 60 or not np.isfinite(inst.ref["diameter"]):
            continue
        d = orc.diameter_continuous_f64(inst.tables.T)
        assert rel(d, inst.ref["diameter"]) < 2e-3 and rel(d, inst.ref["cached_diameter"]) < 2e-3, inst.name
        n += 1
    assert n >= 5


@pytest.mark.gpu
def test_suite_hardness_against_reference(suite):
    """every instance through hardness_of_instance on the GPU: gaps and value norm against the reference's own
    numbers, diameter against the reference / its cache files.  The reference's Q and V are early-stopped iterates
    (eps = 1e-3, Gauss-Seidel), ours sit at the fixed point: near-tied actions move the regularised gaps
    1/(V-Q+0.1) by up to ~1e-2 relative (SimpleGrid), the value norm by a few 1e-3, the diameter by < 1e-3"""
    from colosseum_b200.suite import hardness_of_instance

    checked = {"gaps": 0, "value_norm": 0, "diameter": 0}
    for inst in suite:
        big = inst.S * max(inst.H, 1) > 4000 or inst.S > 420
        res = hardness_of_instance(inst, diameter=not big)
        ref = inst.ref
        if np.isfinite(ref["gaps"]):
            assert rel(res["gaps"], ref["gaps"]) < 1.5e-2, (inst.name, res["gaps"], ref["gaps"])
            checked["gaps"] += 1
        if np.isfinite(ref["value_norm"]) and np.isfinite(res["value_norm"]):
            assert abs(res["value_norm"] - ref["value_norm"]) < 5e-3 * max(ref["value_norm"], 0.2), \
                (inst.name, res["value_norm"], ref["value_norm"])
            checked["value_norm"] += 1
        if "diameter" in res:
            for key in ("diameter", "cached_diameter"):
                if np.isfinite(ref[key]):
                    assert rel(res["diameter"], ref[key]) < 2e-3, (inst.name, key, res["diameter"], ref[key])
                    checked["diameter"] += 1
    assert checked["gaps"] >= 70 and checked["value_norm"] >= 50 and checked["diameter"] >= 40, checked


@pytest.mark.gpu
def test_suite_step_phase(suite):
    """the step phase of a C3 work item on one instance per family: visitation totals are conserved and the
    empirical next-state frequencies follow T (chi-square on the most visited (s,a) pair)"""
    import scipy.stats

    from colosseum_b200.batched_mdp import BatchedMDP

    seen = set()
    for inst in suite:
        fam = inst.name.split(".")[0]
        if fam in seen or inst.S > 300:
            continue
        seen.add(fam)
        N, n_steps = 2048, 60
        env = BatchedMDP(inst.tables, N, mode="succ", seed=5)
        env.reset()
        s_prev, a_all, s_next = [], [], []
        for _ in range(n_steps):
            s0 = env.state.clone()
            st0 = env.step_type.clone()
            env.step_async(None, auto_reset=True)
            keep = (st0 != 2).cpu().numpy()  # envs that took a regular step (not the auto-reset path)
            s_prev.append(s0.cpu().numpy()[keep]); a_all.append(env.action.cpu().numpy()[keep])
            s_next.append(env.state.cpu().numpy()[keep])
        assert int(env.visits_s.sum()) == N * (n_steps + 1)
        sp, aa, sn = map(np.concatenate, (s_prev, a_all, s_next))
        key = sp * inst.A + aa
        top = np.bincount(key).argmax()
        s, a = divmod(int(top), inst.A)
        sel = key == top
        p = inst.tables.T[s, a].astype(np.float64)
        cnt = np.bincount(sn[sel], minlength=inst.S)
        assert cnt[p == 0].sum() == 0, inst.name  # zero-probability states are never produced
        live = p > 0
        if live.sum() > 1 and sel.sum() * p[live].min() >= 5:
            chi2 = ((cnt[live] - sel.sum() * p[live] / p[live].sum()) ** 2 / (sel.sum() * p[live] / p[live].sum())).sum()
            assert scipy.stats.chi2.sf(chi2, live.sum() - 1) > 1e-5, (inst.name, chi2)
    assert len(seen) >= 10


@pytest.mark.gpu
def test_concurrent_instances_match_sequential(suite):
    """run_many: several instances in flight (host threads, one stream each) == the same instances one after another"""
    from colosseum_b200.suite import run_many

    work = [(suite[i], 100 + i) for i in (1, 12, 19, 40, 42, 53, 54, 62)]
    seq = run_many(work, n_workers=1, n_envs=256, n_steps=50)
    par = run_many(work, n_workers=4, n_envs=256, n_steps=50)
    for (a, _), (b, _) in zip(seq, par):
        for k in ("gaps", "value_norm", "diameter", "visits_total", "mean_reward_last_step"):
            assert a[k] == b[k] or (np.isnan(a[k]) and np.isnan(b[k])), k


@pytest.mark.gpu
def test_suite_reference_iterates(suite):
    """the reference's OWN early-stopped iterates (sweep_order='gauss_seidel', reference_iterates=True) on the
    continuous benchmark instances: measures computed from them sit on the reference's recorded numbers -- value norm
    to 5e-4, diameter to 2e-3 ABSOLUTE (bit-identical on several), gaps to 1e-4 except where exactly tied actions make
    1/(V-Q+0.1) hypersensitive (SimpleGrid: 6e-3) -- an order of magnitude closer than the fixed-point comparison of
    test_suite_hardness_against_reference, as it must be"""
    import torch

    import colosseum_b200.dynamic_programming as dp
    import colosseum_b200.hardness as hd

    n = 0
    for inst in suite:
        if inst.episodic or inst.S > 420:
            continue
        T = torch.from_numpy(inst.tables.T).cuda()
        R = torch.from_numpy(inst.R).cuda()
        Q, V = dp.discounted_value_iteration(T, R, sweep_order="gauss_seidel")  # reference defaults: 0.99, 1e-3, f32
        ref = inst.ref
        gaps = hd.get_sum_reciprocals_suboptimality_gaps(Q, V)
        assert rel(gaps, ref["gaps"]) < (6e-3 if "SimpleGrid" in inst.name else 5e-4), (inst.name, gaps, ref["gaps"])
        det = bool((inst.tables.succ_len == 1).all()) and all(k == "deterministic" for k, _ in inst.tables.rew_kinds)
        vn = 0.0 if det else hd.calculate_norm_discounted(T, V, precision="f32")
        assert abs(vn - ref["value_norm"]) < 5e-4 * max(ref["value_norm"], 0.05), (inst.name, vn, ref["value_norm"])
        if np.isfinite(ref["diameter"]):
            d = hd.get_diameter(T, False, reference_iterates=True)
            assert abs(d - ref["diameter"]) < 2e-3, (inst.name, d, ref["diameter"])
        n += 1
    assert n >= 18


@pytest.mark.gpu
def test_native_runner_equals_the_python_path(suite):
    """colo_suite_run (C++ worker threads, one stream each) against suite.run_instance (the same C-ABI calls issued from
    Python) on a mixed set of instances: the same numbers -- the walks bit for bit (same Philox counters), the measures
    to rounding of identical kernels -- and the reference's recorded answers"""
    from colosseum_b200.suite import run_instance, run_many_native

    small = [i for i in suite if i.S <= 260]
    pick = [i for i in small if not i.episodic][:8] + [i for i in small if i.episodic][:8]
    assert any(i.episodic for i in pick) and any(not i.episodic for i in pick)
    work = [(inst, 3) for inst in pick]
    native = run_many_native(work, n_workers=4, n_envs=256, n_steps=50)
    for (inst, seed), (res, tm) in zip(work, native):
        ref, _ = run_instance(inst, n_envs=256, n_steps=50, seed=seed)
        assert res["visits_total"] == ref["visits_total"] == 256 * 51
        assert abs(res["mean_reward_last_step"] - ref["mean_reward_last_step"]) < 1e-6
        for k in ("gaps", "value_norm", "diameter"):
            a, b = res[k], ref[k]
            assert (np.isnan(a) and np.isnan(b)) or abs(a - b) <= 1e-9 * max(1.0, abs(b)), (inst.name, k, a, b)
        if not np.isnan(inst.ref["diameter"]):
            assert abs(res["diameter"] - inst.ref["diameter"]) < 2e-3 * inst.ref["diameter"], inst.name
        assert tm["step_s"] > 0 and tm["hardness_s"] > 0


def test_seed_fixture_rebuilds_the_reference_T():
    """tests/golden/c3_suite_seeds.npz (the 14 quick-test sets at seed 0 + all 94 parameter sets at seeds 1..10): the
    dense T rebuilt from the successor lists is bit-identical to the reference's mdp.T (CRC) -- every 12th instance"""
    from colosseum_b200.suite import load_suite_all, suite_size

    n = suite_size(GOLDEN)
    assert n == 80 + 954
    idx = list(range(80, n, 12))
    insts = load_suite_all(GOLDEN, indices=idx)
    assert len(insts) == len(idx) and all(i.check_T() for i in insts)
    assert any(".s10" in i.name for i in insts) and any(i.episodic for i in insts)
    # the continuous classes carry the reference's own cached hardness values at every seed
    assert sum(np.isfinite(i.ref["cached_diameter"]) for i in insts) >= 20
