"""PSRL per-episode parts: Dirichlet sample (f64 / fast), NIG sample, batched episodic VI."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import load_instance
from colosseum_b200.tables import MDPTables
from colosseum_b200 import _cabi
import colosseum_b200.agent_loop as al
import colosseum_b200.dynamic_programming as dp

def timeit(f, n=10):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

for inst, n in (("taxi_epi", 1024), ("deepsea8_epi", 8192), ("c1_riverswim_epi", 8192)):
    tb = MDPTables.from_golden(load_instance(inst))
    ag = al.PSRLEpisodic(0, tb, 10 ** 6, n_loops=n)
    ag.steps(3 * tb.H)
    lib = _cabi.lib(); st = _cabi.current_stream()
    rows = n * tb.S * tb.A
    t_f64 = timeit(lambda: lib.colo_sample_dirichlet_rows(ag.dir_hyper.data_ptr(), rows, tb.S, 0, 1, 0, ag.T_sample.data_ptr(), st))
    t_fast = timeit(lambda: lib.colo_sample_dirichlet_rows_fast(ag.dir_hyper.data_ptr(), rows, tb.S, 0, 1, 0, ag.T_sample.data_ptr(), st))
    t_nig = timeit(lambda: lib.colo_sample_nig_rewards(ag.nig_hyper.data_ptr(), rows, 0, 1, 0, ag.R_sample.data_ptr(), st))
    t_vi = timeit(lambda: dp.episodic_value_iteration(tb.H, ag.T_sample, ag.R_sample, precision="f32"))
    t_steps = timeit(lambda: _cabi.lib().colo_psrl_episodic_steps(__import__("ctypes").byref(ag.dev.c), __import__("ctypes").byref(ag._args), tb.H, 10, st))
    gb = rows * tb.S * 4 / 1e9
    print(f"{inst:18s} loops={n}: T = {gb:.2f} GB; dirichlet f64 {t_f64:8.1f} us, fast {t_fast:8.1f} us ({rows * tb.S / t_fast / 1e3:.1f} G draws/s), "
          f"nig {t_nig:6.1f} us, episodic VI {t_vi:8.1f} us ({gb * tb.H / t_vi * 1e6 / 1e3:.0f} GB/s of T), steps {t_steps:6.1f} us", flush=True)
