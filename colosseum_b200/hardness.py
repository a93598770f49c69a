"""B200 twin of `colosseum.hardness.measures` (colosseum/hardness/measures/__init__.py:4-12).

  get_diameter                               hardness/measures/diameter.py:20-39
  calculate_norm_discounted                  hardness/measures/value_norm.py:85-87
  get_sum_reciprocals_suboptimality_gaps     hardness/measures/sum_reciprocals_suboptimality_gaps.py:6-28

The diameter is the dominant hardness cost of the reference (S independent VI-like solves, seconds to minutes per
MDP).  Here all targets are iterated at once: every sweep is one launch of the multi-target hitting-time backup
(T shared by all targets and re-served from L2), each target frozen on the device as soon as it has converged.
Hardness measures are scalars, so they are computed with fp64 accumulation unless precision='f32' is requested.
"""
import ctypes as C

import numpy as np

from . import _cabi
from .dynamic_programming import _is_tensor, _result, _scratch, _torch, to_device


def get_diameter(T, is_episodic, max_value=None, *, precision="f64", epsilon=None, targets=None,
                 max_iter=int(1e6), return_sweeps=False, reference_iterates=None):
    """colosseum/hardness/measures/diameter.py:20-39.  T is [S,A,S] (continuous) or the episodic [H,S,A,S] tensor of
    mdp/utils/mdp_creation.py:98-128.  Returns the diameter, or None when a hitting time exceeds max_value.

    The reference stops each target at eps=1e-3 with order-dependent early exits, which makes its own value
    path-dependent at ~1e-5 relative; the GPU iterates every target to `epsilon`
    (default 1e-9 in f64, 1e-4 in f32) on the sweep-to-sweep change instead."""
    Td = to_device(T)
    assert (is_episodic and Td.dim() == 4) or (not is_episodic and Td.dim() == 3)
    torch = _torch()
    if reference_iterates is None:
        from .dynamic_programming import get_sweep_order

        reference_iterates = get_sweep_order() == "gauss_seidel"
    ref_it = bool(reference_iterates)
    gs = ref_it and not is_episodic
    if ref_it and precision == "f64" and epsilon is None:
        precision = "f32"  # the reference's arithmetic type: its own iterate is a float32 one
    f64 = precision == "f64"
    # reference_iterates: the reference's own iterate at ITS epsilon = 1e-3.  Continuous MDPs: per-target in-place VI
    # (diameter.py:91 -> infinite_horizon.py:121-142).  Episodic MDPs: the layered kernels already sweep like
    # _episodic_diameter_calculation (diameter.py:285-318: layer h-1 from the freshly updated layer h), so float32 and
    # the reference's epsilon are all it takes (its order-dependent loose exit, :316, is not reproduced).
    eps = float(epsilon if epsilon is not None else (1e-3 if ref_it else (1e-9 if f64 else 1e-4)))
    lib = _cabi.lib()
    S, A = int(Td.shape[-1]), int(Td.shape[-2])
    tg = np.arange(S, dtype=np.int32) if targets is None else np.ascontiguousarray(targets, np.int32)
    K = len(tg)
    tgd = torch.from_numpy(tg).cuda()
    out = (C.c_double * 2)()
    mv = float(max_value) if max_value is not None else 0.0
    if is_episodic:
        H = int(Td.shape[0])
        work = _scratch(lib.colo_diameter_episodic_work_bytes(K, H, S, A, int(f64)))
        fn = lib.colo_diameter_episodic_f64acc if f64 else lib.colo_diameter_episodic_f32
        rc = fn(_cabi.ptr(Td), _cabi.ptr(tgd), K, H, S, A, eps, mv, int(max_iter), _cabi.ptr(work), out,
                _cabi.current_stream())
    elif gs:
        work = _scratch(lib.colo_diameter_continuous_gs_work_bytes(K, S, int(f64)))
        fn = lib.colo_diameter_continuous_gs_f64acc if f64 else lib.colo_diameter_continuous_gs_f32
        rc = fn(_cabi.ptr(Td), _cabi.ptr(tgd), K, S, A, eps, mv, int(max_iter), _cabi.ptr(work), out,
                _cabi.current_stream())
    else:
        work = _scratch(lib.colo_diameter_continuous_work_bytes(K, S, int(f64)))
        fn = lib.colo_diameter_continuous_f64acc if f64 else lib.colo_diameter_continuous_f32
        rc = fn(_cabi.ptr(Td), _cabi.ptr(tgd), K, S, A, eps, mv, int(max_iter), _cabi.ptr(work), out,
                _cabi.current_stream())
    _cabi.check(rc, "colo_diameter")
    if rc == _cabi.OVERFLOW:
        return None
    if rc == _cabi.MAX_ITER:
        from .dynamic_programming import DynamicProgrammingMaxIterationExceeded

        raise DynamicProgrammingMaxIterationExceeded()
    d = float(out[0])
    if max_value is not None and d > max_value:
        return None
    return (d, int(out[1])) if return_sweeps else d


def calculate_norm_discounted(T, V, *, precision="f64"):
    """colosseum/hardness/measures/value_norm.py:85-87 (Ev is indexed by the next state, as in the reference)."""
    torch = _torch()
    f64 = precision == "f64"
    Td = to_device(T)
    Vd = to_device(V, np.float64 if f64 else np.float32).reshape(-1)
    S, A, _ = Td.shape
    lib = _cabi.lib()
    work = _scratch(lib.colo_value_norm_work_bytes(S, A, int(f64)))
    out = torch.zeros(1, dtype=torch.float64 if f64 else torch.float32, device="cuda")
    fn = lib.colo_value_norm_f64acc if f64 else lib.colo_value_norm_f32
    rc = fn(_cabi.ptr(Td), _cabi.ptr(Vd), S, A, _cabi.ptr(work), _cabi.ptr(out), _cabi.current_stream())
    _cabi.check(rc, "colo_value_norm")
    return float(out.item())


def calculate_norm_average(T, tps, average_rewards, steps=1000):
    """colosseum/hardness/measures/value_norm.py:90-93 (with _calculate_bias :69-82): the undiscounted environmental
    value norm, i.e. the norm of :85-87 taken of the bias vector h = sum_{i<steps} P^i (r - P^steps r)."""
    torch = _torch()
    Td, Pd = to_device(T), to_device(tps)
    rd = to_device(average_rewards).reshape(-1)
    S, A, _ = Td.shape
    lib = _cabi.lib()
    h = torch.empty(S, dtype=torch.float64, device="cuda")
    work = _scratch(lib.colo_bias_series_work_bytes(S))
    rc = lib.colo_bias_series_f64(_cabi.ptr(Pd), _cabi.ptr(rd), S, int(steps), _cabi.ptr(h), _cabi.ptr(work),
                                  _cabi.current_stream())
    _cabi.check(rc, "colo_bias_series_f64")
    return calculate_norm_discounted(Td, h, precision="f64")


def get_sum_reciprocals_suboptimality_gaps(Q, V, reachable_states=None, regularization=0.1):
    """colosseum/hardness/measures/sum_reciprocals_suboptimality_gaps.py:6-28."""
    torch = _torch()
    Qd = to_device(Q, np.float64)
    Vd = to_device(V, np.float64)
    is_episodic = Vd.dim() == 2
    mask = None
    if is_episodic:
        assert reachable_states is not None, (
            "For the episodic setting, it is necessary to provide the set of nodes that are reachable for any given"
            "in episode time step."
        )
        m = np.zeros(tuple(Vd.shape), np.uint8)
        # the reference stacks gaps[h, s] once per listed pair (duplicates would count twice; the list has none)
        hs = (reachable_states if isinstance(reachable_states, np.ndarray)
              else np.asarray(list(reachable_states), np.int64)).reshape(-1, 2)
        m[hs[:, 0], hs[:, 1]] = 1
        assert m.sum() == len(hs), "duplicate (h, s) pairs in reachable_states"
        mask = torch.from_numpy(m).cuda()
    A = int(Qd.shape[-1])
    NS = int(Vd.numel())
    out = torch.zeros(1, dtype=torch.float64, device="cuda")
    rc = _cabi.lib().colo_gaps_f64(_cabi.ptr(Qd), _cabi.ptr(Vd), _cabi.ptr(mask), NS, A, float(regularization),
                                   _cabi.ptr(out), _cabi.current_stream())
    _cabi.check(rc, "colo_gaps_f64")
    return float(out.item())
