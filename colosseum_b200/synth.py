"""Synthetic dense MDPs generated on the device (BASELINE.json config C5: S=40,000, A=8, ~51 GB of T)."""
from . import _cabi


def synth_dense_rows(row0, nrows, S, A, seed=0):
    """rows [row0, row0+nrows) of the synthetic MDP `seed`: (T_rows f32[nrows,A,S], R_rows f32[nrows,A]) on the
    current CUDA device; a pure function of (seed, global row), hence identical for every sharding."""
    import torch

    _cabi.require_cuda()
    T = torch.empty((nrows, A, S), dtype=torch.float32, device="cuda")
    R = torch.empty((nrows, A), dtype=torch.float32, device="cuda")
    rc = _cabi.lib().colo_synth_dense_rows(_cabi.ptr(T), _cabi.ptr(R), int(row0), int(nrows), int(S), int(A), int(seed),
                                           _cabi.current_stream())
    _cabi.check(rc, "colo_synth_dense_rows")
    return T, R
