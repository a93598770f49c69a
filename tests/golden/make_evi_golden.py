"""Generate tests/golden/evi.npz by running the UNMODIFIED reference's numba `extended_value_iteration`
(colosseum/dynamic_programming/infinite_horizon.py:67-118) on UCRL2-shaped inputs: an empirical model P = N / N.sum
from simulated visit counts, Chernoff confidence widths as colosseum/agent/agents/infinite_horizon/ucrl2.py:240-296
computes them (beta_p is [S,A,1] float64), float32 estimated rewards.  Build container only (needs /root/reference)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.reference_import import import_reference  # noqa: E402

import_reference()
from colosseum.dynamic_programming.infinite_horizon import extended_value_iteration  # noqa: E402


def chernoff(it, N, delta, sqrt_C, log_C, rng=1.0):  # ucrl2.py:22-24
    return rng * np.sqrt(sqrt_C * np.log(log_C * (it + 1) / delta) / np.maximum(1, N))


def case(seed, S, A, n_obs, succ, alpha=0.5):
    rs = np.random.RandomState(seed)
    # a sparse "true" MDP with `succ` successors per (s,a), then counts from n_obs simulated transitions per pair
    Ttrue = np.zeros((S, A, S))
    for s in range(S):
        for a in range(A):
            js = rs.choice(S, size=succ, replace=False)
            Ttrue[s, a, js] = rs.dirichlet(np.ones(succ))
    N = np.stack([[rs.multinomial(rs.randint(1, n_obs), Ttrue[s, a]) for a in range(A)] for s in range(S)]).astype(np.float64)
    P = (N / N.sum(-1, keepdims=True)).astype(np.float32)
    nb = N.sum(-1)
    it = int(nb.sum())
    est = (rs.uniform(0, 1, (S, A)) ** 3).astype(np.float32)
    beta_r = alpha * chernoff(it, nb, 0.05, 3.5, 2 * S * A)
    beta_p = alpha * chernoff(it, nb, 0.05, 14 * S, 2 * A).reshape(S, A, 1)
    return P, est, beta_r, beta_p


def main():
    out = {}
    for i, (seed, S, A, n_obs, succ, alpha) in enumerate([(0, 12, 2, 4000, 3, 0.05), (1, 40, 3, 20000, 4, 0.02), (2, 25, 4, 3000, 2, 0.5),
                                                         (3, 64, 2, 100000, 5, 0.01)]):
        P, est, beta_r, beta_p = case(seed, S, A, n_obs, succ, alpha)
        for eps in (1e-3, 1e-5):
            span, Q, V = extended_value_iteration(P, est, beta_r, beta_p, 1.0, eps)
            tag = f"{i}_{'loose' if eps == 1e-3 else 'tight'}"
            out.update({f"span_{tag}": span, f"Q_{tag}": Q, f"V_{tag}": V})
            print(f"case {i} S={S} A={A} eps={eps}: span={span:.6f} V[:3]={V[:3]}")
        out.update({f"P_{i}": P, f"est_{i}": est, f"beta_r_{i}": beta_r, f"beta_p_{i}": beta_p})
    out["n_cases"] = 4
    np.savez_compressed(os.path.join(HERE, "evi.npz"), **out)


if __name__ == "__main__":
    main()
