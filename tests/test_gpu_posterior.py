"""GPU: PSRL's Dirichlet model sample (conjugate_transitions.py:48-60) -- distributional parity with numpy/scipy."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("fast", [False, True])
def test_dirichlet_rows_distribution(fast):
    import functools

    import scipy.stats

    from colosseum_b200.posterior import sample_transition_model

    sample_transition_model = functools.partial(sample_transition_model, fast=fast)
    rs = np.random.RandomState(0)
    S = 6
    alpha = np.array([0.05, 0.5, 1.0, 2.5, 30.0, 400.0], np.float32)
    n = 60000
    T = sample_transition_model(np.tile(alpha, (n, 1)), seed=3, t=0)
    assert T.shape == (n, S) and T.dtype == np.float32
    assert np.all(T >= 0) and np.allclose(T.sum(-1), 1.0, atol=2e-5)  # 1e-5 in the denominator (sic)
    a0 = float(alpha.sum())
    for j in range(S):
        marg = scipy.stats.beta(float(alpha[j]), a0 - float(alpha[j]))  # Dirichlet marginals are Beta
        x = T[:, j].astype(np.float64)
        assert abs(x.mean() - marg.mean()) < 6 * marg.std() / np.sqrt(n) + 1e-6, j
        # standard error of a sample variance: var * sqrt((kurtosis - 1) / n) -- large for the heavy-tailed alpha << 1
        # components (Beta(0.05, .): kurtosis > 100), so the bound is 5 standard errors where that exceeds 5 %
        kurt = float(marg.stats(moments="k")) + 3.0
        se = marg.var() * np.sqrt(max(kurt - 1.0, 2.0) / n)
        assert abs(x.var() - marg.var()) < max(0.05 * marg.var(), 5 * se) + 1e-9, j
        if alpha[j] >= 0.5:
            assert scipy.stats.kstest(x, marg.cdf).pvalue > 1e-4, j
    # same numpy recipe as the reference, same moments
    r = rs.standard_gamma(np.tile(alpha, (n, 1))).astype(np.float32)
    Tn = r / (1e-5 + r.sum(-1, keepdims=True))
    assert np.abs(T.mean(0) - Tn.mean(0)).max() < 3e-3
    # counter semantics: (seed, t) reproducible, t and row0 move the stream, row sharding == unsharded
    T2 = sample_transition_model(np.tile(alpha, (n, 1)), seed=3, t=0)
    assert np.array_equal(T, T2)
    assert not np.array_equal(T, sample_transition_model(np.tile(alpha, (n, 1)), seed=3, t=1))
    half = sample_transition_model(np.tile(alpha, (n // 2, 1)), seed=3, t=0, row0=n // 2)
    assert np.array_equal(half, T[n // 2:])


def test_sampled_model_feeds_value_iteration():
    """posterior_sampling.py:142-144: episodic_value_iteration(H, *model.sample()) with the model kept on the GPU"""
    import torch

    import colosseum_b200.dynamic_programming as dp
    from colosseum_b200.posterior import sample_transition_model
    from oracle import oracle as orc

    S, A, H = 30, 3, 8
    hyper = torch.full((S, A, S), 0.3, device="cuda")
    T = sample_transition_model(hyper, seed=1, t=5)
    R = torch.rand((S, A), device="cuda")
    Q, V = dp.episodic_value_iteration(H, T, R, precision="f64")
    Qo, Vo = orc.episodic_f64(H, T.cpu().numpy(), R.cpu().numpy())
    np.testing.assert_allclose(V.cpu().numpy(), Vo, rtol=1e-6)
