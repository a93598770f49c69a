"""B200 twin of PSRL's transition-model sample (colosseum/agent/mdp_models/bayesian_models/conjugate_transitions.py:
48-60, `M_DIR.sample`): the S*A*S gamma draws + the reference's normalisation on the GPU, so that the sampled model
never leaves HBM before value iteration runs on it (posterior_sampling.py:142-144: `episodic_value_iteration(H,
*model.sample())`)."""
import numpy as np

from . import _cabi
from .dynamic_programming import _is_tensor, _result, _torch, to_device


def sample_transition_model(hyper_params, seed=0, t=0, row0=0, fast=False):
    """hyper_params: Dirichlet parameters [S,A,S] (or rows [n,S]); returns T with the same shape, float32
    (`r / (1e-5 + r.sum(-1))`, sic).  `t` is the draw counter: same (seed, t) -> same sample.  fast=True: single-precision
    gamma draws (same distribution to ~1e-6, ~20x the throughput)."""
    torch = _torch()
    as_numpy = not _is_tensor(hyper_params)
    h = to_device(hyper_params)
    S = int(h.shape[-1])
    rows = int(h.numel() // S)
    T = torch.empty_like(h)
    fn = _cabi.lib().colo_sample_dirichlet_rows_fast if fast else _cabi.lib().colo_sample_dirichlet_rows
    rc = fn(_cabi.ptr(h), rows, S, int(row0), int(seed), int(t), _cabi.ptr(T),
                                                _cabi.current_stream())
    _cabi.check(rc, "colo_sample_dirichlet_rows")
    return _result(T, as_numpy)
