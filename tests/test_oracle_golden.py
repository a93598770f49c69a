"""Pins the CPU oracle (oracle/) to the unmodified reference: every oracle function is checked against golden
vectors recorded from the reference by tests/golden/make_golden.py, the reference's cached hardness files and the
numbers printed in its executed notebooks.  CPU only."""
import json
import os

import numpy as np
import pytest

from conftest import CONTINUOUS, EPISODIC, GOLDEN, load_instance
from colosseum_b200.tables import MDPTables
from oracle import oracle as orc


def rel(a, b):
    return abs(a - b) / max(abs(b), 1e-12)


# ------------------------------------------------------------------------------------------------ sampler
def test_inverse_cdf_matches_cpython_choices(sampler_kat):
    """succ mode and dense-f64 mode == random.Random(seed).choices (custom_samplers.py:49-72), 5 x 5000 draws."""
    for i in range(5):
        probs, us, chosen = sampler_kat[f"probs_{i}"], sampler_kat[f"u_{i}"], sampler_kat[f"chosen_{i}"]
        n = len(probs)
        tb = MDPTables.from_successors(n, 1, np.tile(np.arange(n, dtype=np.int32), (n, 1, 1)),
                                       np.tile(probs, (n, 1, 1)), np.full((n, 1), n), np.zeros((n, 1, n), np.int32),
                                       [("deterministic", (0.0,))], [0], [1.0])
        N = len(us)
        for mode in (2, 1):
            ht = orc.HostTables(n, 1, succ_cum=tb.succ_cum, succ_idx=tb.succ_idx, succ_len=tb.succ_len,
                                rew_cls_succ=tb.rew_cls_succ, rew_q=tb.rew_q,
                                cdf=orc.build_dense_cdf(tb.T, f64=True) if mode == 1 else None)
            if mode == 1:
                # dense rows must carry the sampler's fp64 probabilities to be bit-exact: rebuild from them
                cdf = np.zeros((n, 1, tb.ld))
                cdf[:, 0, :n] = np.cumsum(probs)
                cdf[:, 0, n:] = cdf[:, 0, n - 1:n]
                ht = orc.HostTables(n, 1, cdf=cdf, rew_q=tb.rew_q)
            state = np.zeros(N, np.int32)
            h = np.zeros(N, np.int32)
            st = np.ones(N, np.uint8)
            _, obs, rc, _ = orc.env_step(ht, mode, state, h, st, action=np.zeros(N, np.int32), u_next=us,
                                         u_rew=np.zeros(N, np.float32))
            assert rc == 0
            assert (state == chosen).all(), f"kat {i} mode {mode}"


def test_philox_known_answer():
    """Philox4x32-10 known-answer vectors of Random123 (kat_vectors: zero and ff.. inputs)."""
    assert orc.philox(0, 0, 0) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    got = orc.philox(0xFFFFFFFFFFFFFFFF, 0xFFFFFFFFFFFFFFFF, 0xFFFFFFFFFFFFFFFF)
    assert got == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]


# ------------------------------------------------------------------------------------------------ DP
def test_dp_synth(dp_synth):
    g = dp_synth
    for b in range(3):
        T, R, pi = g[f"T_{b}"], g[f"R_{b}"], g[f"pi_{b}"]
        # the reference iterate (fp32 Gauss-Seidel, eps=1e-3): same iterates up to BLAS summation order
        Q, V, _ = orc.discounted_gs_f32(T, R, gamma=0.99, eps=1e-3)
        np.testing.assert_allclose(V, g[f"V_{b}"], rtol=3e-5, atol=2e-4)
        np.testing.assert_allclose(Q, g[f"Q_{b}"], rtol=3e-5, atol=2e-4)
        # fixed point (fp64) vs the reference run to eps=1e-6: |V* - V_ref| <= eps*gamma/(1-gamma) + fp32 noise
        Q64, V64, _ = orc.discounted_f64(T, R, gamma=float(np.float32(0.99)), tol=1e-13)
        np.testing.assert_allclose(V64, g[f"Vt_{b}"], rtol=1e-5, atol=2e-4)
        np.testing.assert_allclose(Q64, g[f"Qt_{b}"], rtol=1e-5, atol=2e-4)
        Qg, Vg, _ = orc.discounted_f64(T, R, gamma=float(np.float32(0.99)), tol=1e-13, gauss_seidel=True)
        np.testing.assert_allclose(Vg, V64, rtol=1e-10)
        # policy evaluation (reference eps=1e-7)
        Qp, Vp, _ = orc.discounted_f64(T, R, pi=pi, gamma=float(np.float32(0.99)), tol=1e-13)
        np.testing.assert_allclose(Vp, g[f"Vp_{b}"], rtol=2e-5)
        np.testing.assert_allclose(Qp, g[f"Qp_{b}"], rtol=2e-5)
        # episodic VI / PE are exact recurrences: only fp32 vs fp64 summation differs
        H = int(g[f"H_{b}"])
        Qe, Ve = orc.episodic_f64(H, T, R)
        np.testing.assert_allclose(Ve, g[f"Ve_{b}"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(Qe, g[f"Qe_{b}"], rtol=1e-5, atol=1e-6)
        Qe32, Ve32 = orc.episodic_f32(H, T, R)
        np.testing.assert_allclose(Ve32, g[f"Ve_{b}"], rtol=1e-5, atol=1e-6)
        Qpe, Vpe = orc.episodic_f64(H, T, R, pi=g[f"pol_{b}"])
        np.testing.assert_allclose(Vpe, g[f"Vpe_{b}"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(Qpe, g[f"Qpe_{b}"], rtol=1e-5, atol=1e-6)
        # hardness on the same MDP
        # diameter: fixed point vs the reference's per-target VI run to eps=2e-5; the reference's default eps=1e-3
        # stops early (gamma=1: the early-stop error is not bounded by eps), hence the looser second bound
        d = orc.diameter_continuous_f64(T)
        assert rel(d, float(g[f"diam_tight_{b}"])) < 2e-5
        assert rel(d, float(g[f"diam_{b}"])) < 2e-3
        assert rel(orc.value_norm_f64(T, g[f"V_{b}"]), float(g[f"vnorm_{b}"])) < 1e-5
        assert rel(orc.gaps_f64(g[f"Q_{b}"], g[f"V_{b}"]), float(g[f"gaps_{b}"])) < 1e-5
        # overflow contract: None (infinite_horizon.py:136-138, finite_horizon.py:24-25)
        assert orc.discounted_gs_f32(T, R, max_abs=5.0) is None
        assert orc.discounted_f64(T, R, max_abs=5.0) is None
        assert orc.episodic_f64(H, T, R, max_value=1.5) is None


def test_c1_anchor_and_doc_goldens():
    doc = json.load(open(os.path.join(GOLDEN, "doc_goldens.json")))
    g = load_instance("c1_riverswim_epi")
    Q, V = orc.episodic_f64(int(g["H"]), g["T"], g["R"])
    np.testing.assert_allclose(V[0], doc["c1_V0"], rtol=2e-6)
    np.testing.assert_allclose(V, g["vi_V"], rtol=2e-6, atol=1e-7)
    g = load_instance("doc_simplegrid4")
    assert rel(orc.diameter_continuous_f64(g["T"]), doc["diameter"]) < 2e-3
    assert rel(orc.diameter_continuous_f64(g["T"]), float(g["diameter_tight"])) < 2e-5
    Q, V, _ = orc.discounted_f64(g["T"], g["R"], gamma=float(np.float32(0.99)), tol=1e-13)
    # the notebook numbers come from the reference's early-stopped (eps=1e-3) V: looser on the end-to-end values
    assert rel(orc.value_norm_f64(g["T"], V), doc["value_norm"]) < 2e-3
    assert rel(orc.gaps_f64(Q, V), doc["suboptimal_gaps"]) < 2e-3
    assert rel(orc.value_norm_f64(g["T"], g["vi_V"]), doc["value_norm"]) < 1e-5
    assert rel(orc.gaps_f64(g["vi_Q"], g["vi_V"]), doc["suboptimal_gaps"]) < 1e-5


@pytest.mark.parametrize("name", CONTINUOUS)
def test_continuous_instances(name):
    g = load_instance(name)
    T, R = g["T"], g["R"]
    S, A = R.shape
    gam = float(np.float32(0.99))
    Q, V, _ = orc.discounted_f64(T, R, gamma=gam, tol=1e-13)
    np.testing.assert_allclose(V, g["vi_tight_V"], rtol=1e-5, atol=3e-4)
    np.testing.assert_allclose(Q, g["vi_tight_Q"], rtol=1e-5, atol=3e-4)
    # reference default eps=1e-3 stops within eps*gamma/(1-gamma) ~ 0.1 of the fixed point
    assert np.abs(V - g["vi_V"]).max() < 0.11
    pi = np.ones((S, A), np.float32) / A
    Qp, Vp, _ = orc.discounted_f64(T, R, pi=pi, gamma=gam, tol=1e-13)
    np.testing.assert_allclose(Vp, g["pe_V"], rtol=5e-5, atol=1e-5)
    # measures, formula-level (reference inputs) ...
    if not bool(g["all_deterministic"]):
        assert rel(orc.value_norm_f64(T, g["vi_V"]), float(g["value_norm"])) < 5e-5  # reference einsum is fp32
        cached = float(g["cached_value_norm"])
        if np.isfinite(cached):
            assert rel(orc.value_norm_f64(T, g["vi_V"]), cached) < 1e-4
    assert rel(orc.gaps_f64(g["vi_Q"], g["vi_V"]), float(g["gaps"])) < 1e-5
    # ... and the diameter end to end (fixed point vs the reference's early-exit value and its cache file)
    if S <= 110:
        d = orc.diameter_continuous_f64(T)
        assert rel(d, float(g["diameter_tight"])) < 2e-5  # reference VI kernel at eps=2e-5
        assert rel(d, float(g["diameter"])) < 2e-3        # reference default (early stop at eps=1e-3)
        cached = float(g["cached_diameter"])
        if np.isfinite(cached):
            assert rel(d, cached) < 2e-3


@pytest.mark.parametrize("name", EPISODIC)
def test_episodic_instances(name):
    g = load_instance(name)
    T, R, H = g["T"], g["R"], int(g["H"])
    S, A = R.shape
    Q, V = orc.episodic_f64(H, T, R)
    np.testing.assert_allclose(V, g["vi_V"], rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(Q, g["vi_Q"], rtol=2e-5, atol=1e-6)
    pol = np.ones((H, S, A), np.float32) / A
    Qp, Vp = orc.episodic_f64(H, T, R, pi=pol)
    np.testing.assert_allclose(Vp, g["pe_V"], rtol=2e-5, atol=1e-6)
    # gaps over the reachable (h,s) pairs
    mask = np.zeros((H + 1, S), np.uint8)
    mask[g["reach_h"], g["reach_s"]] = 1
    assert rel(orc.gaps_f64(g["vi_Q"], g["vi_V"], mask), float(g["gaps"])) < 1e-5
    # value norm on the continuous form (mdp/base.py:1049-1056)
    assert rel(orc.value_norm_f64(g["T_cf"], g["vi_cf_V"]), float(g["value_norm"])) < 1e-5
    # continuous-form VI (base_finite.py:167-178): fixed point within the reference's eps*gamma/(1-gamma)
    if g["T_cf"].shape[0] <= 700:
        Qc, Vc, _ = orc.discounted_f64(g["T_cf"], g["R_cf"], gamma=float(np.float32(0.99)), tol=1e-12)
        assert np.abs(Vc - g["vi_cf_V"]).max() < 0.11
    # episodic diameter in the augmented space (diameter.py:285-318); reference stops at eps=1e-3 / early exits
    if S * H <= 1500:
        d = orc.diameter_episodic_f64(g["T_epi"])
        assert rel(d, float(g["diameter"])) < 2e-3


# ------------------------------------------------------------------------------------------------ step
@pytest.mark.parametrize("name", CONTINUOUS + EPISODIC)
def test_reference_trajectory_replay(name):
    """BaseMDP.reset/step with auto_reset (base.py:1268-1317) replayed by the oracle from the recorded actions and
    the uniforms the reference's samplers consumed: observations, step types, h-driven termination, visitation
    counts -- and rewards when they are deterministic -- are identical."""
    g = load_instance(name)
    tb = MDPTables.from_golden(g)
    ht = orc.HostTables(tb.S, tb.A, H=tb.H, succ_cum=tb.succ_cum, succ_idx=tb.succ_idx, succ_len=tb.succ_len,
                        rew_cls_succ=tb.rew_cls_succ, rew_q=tb.rew_q, rmin=tb.rmin, rmax=tb.rmax,
                        start_cum=tb.start_cum, start_idx=tb.start_idx)
    acts, us = g["traj_action"], np.nan_to_num(g["traj_u"], nan=0.5)
    vis_s = np.zeros(tb.S, np.uint64)
    vis_sa = np.zeros((tb.S, tb.A), np.uint64)
    state, h, st, obs = orc.env_reset(ht, 1, u_next=us[:1], visits_s=vis_s)
    assert int(st[0]) == int(g["traj_step_type"][0]) and int(obs[0]) == int(g["traj_obs"][0])
    deterministic_rewards = all(k == "deterministic" for k, _ in tb.rew_kinds)
    for t in range(1, len(acts)):
        r, obs, rc, _ = orc.env_step(ht, 2, state, h, st, action=acts[t:t + 1].copy(), u_next=us[t:t + 1],
                                     u_rew=np.zeros(1, np.float32), auto_reset=True, visits_s=vis_s, visits_sa=vis_sa)
        assert rc == 0
        assert int(st[0]) == int(g["traj_step_type"][t]), f"step {t}"
        assert int(obs[0]) == int(g["traj_obs"][t]), f"step {t}"
        if int(st[0]) != 0 and deterministic_rewards:
            assert abs(float(r[0]) - float(g["traj_reward"][t])) < 1e-6
    assert (vis_s.astype(np.int64) == g["traj_visits_s"]).all()
    assert (vis_sa.astype(np.int64) == g["traj_visits_sa"]).all()


def test_step_without_reset_is_flagged():
    g = load_instance("c1_riverswim_epi")
    tb = MDPTables.from_golden(g)
    ht = orc.HostTables(tb.S, tb.A, H=tb.H, succ_cum=tb.succ_cum, succ_idx=tb.succ_idx, succ_len=tb.succ_len,
                        rew_cls_succ=tb.rew_cls_succ, rew_q=tb.rew_q, start_cum=tb.start_cum, start_idx=tb.start_idx)
    state, h, st, obs = orc.env_reset(ht, 4)
    for _ in range(tb.H):
        _, obs, rc, _ = orc.env_step(ht, 2, state, h, st, action=np.zeros(4, np.int32))
        assert rc == 0
    assert (st == 2).all() and (obs == -1).all()
    _, _, rc, _ = orc.env_step(ht, 2, state, h, st, action=np.zeros(4, np.int32))
    assert rc == orc.NEEDS_RESET  # the reference asserts `not self.necessary_reset` (base.py:1291)


def test_reward_tables_match_scipy():
    """the tabulated quantile draw is distributionally the reference's scipy frozen distribution
    (base.py:1196-1207), including the heavy-tailed Beta(0.01, 0.11) of config C1."""
    import scipy.stats

    g = load_instance("c1_riverswim_epi")
    tb = MDPTables.from_golden(g)
    ht = orc.HostTables(tb.S, tb.A, rew_q=tb.rew_q)
    rs = np.random.RandomState(0)
    N = 200000
    for c, (kind, args) in enumerate(tb.rew_kinds):
        u = rs.random_sample(N).astype(np.float32)
        q = tb.rew_q[c]
        t = u * np.float32(len(q) - 1)
        i = np.minimum(t.astype(np.int32), len(q) - 2)
        draw = q[i] + (t - i) * (q[i + 1] - q[i])
        dist = getattr(scipy.stats, kind)(*args)
        # Kolmogorov distance on x >= 1e-30 (below that the fp32 table underflows to 0 where scipy's fp64 draws are
        # 1e-60-like; indistinguishable once rescaled into the rewards range)
        xs = np.concatenate([np.logspace(-30, -1, 300), np.linspace(0.1, 1.0, 300)])
        emp = np.searchsorted(np.sort(draw), xs, side="right") / N
        ks = np.abs(emp - dist.cdf(xs)).max()
        assert ks < 0.005, (kind, args, ks)
        assert abs(draw.mean() - dist.mean()) < 4 * dist.std() / np.sqrt(N) + 1e-4


@pytest.mark.parametrize("name", EPISODIC)
def test_episodic_tensor_forms_restatement(name):
    """oracle restatement of mdp_creation.py:98-176 vs the tensors the reference itself built"""
    g = load_instance(name)
    H = int(g["H"])
    T_epi, R_epi, reach = orc.episodic_T(H, g["T"], g["R"], g["start_idx"], g["start_prob"])
    assert np.array_equal(T_epi, g["T_epi"])
    assert set(zip(*np.nonzero(reach))) == set(zip(g["reach_h"].tolist(), g["reach_s"].tolist()))
    T_cf, R_cf = orc.continuous_form(H, g["T"], g["R"], g["start_idx"], g["start_prob"], zip(g["reach_h"], g["reach_s"]))
    assert np.array_equal(T_cf, g["T_cf"]) and np.array_equal(R_cf, g["R_cf"])


def test_extended_value_iteration_restatement():
    """oracle restatement of UCRL2's extended VI (infinite_horizon.py:67-118, :222-251) vs the reference's numba run"""
    g = np.load(os.path.join(GOLDEN, "evi.npz"))
    for i in range(int(g["n_cases"])):
        for tag, eps in (("loose", 1e-3), ("tight", 1e-5)):
            span, Q, V, it = orc.extended_vi_f32(g[f"P_{i}"], g[f"est_{i}"], g[f"beta_r_{i}"], g[f"beta_p_{i}"], 1.0, eps)
            # same float32 iterates; only the order of the float32 dot product may differ from numba's BLAS call, so
            # the stopping iteration can move by one: differences stay below eps
            assert abs(span - float(g[f"span_{i}_{tag}"])) < 2 * eps + 1e-5, (i, tag)
            np.testing.assert_allclose(Q, g[f"Q_{i}_{tag}"], atol=2 * eps + 2e-6)
            np.testing.assert_allclose(V, g[f"V_{i}_{tag}"], atol=2 * eps + 2e-6)
