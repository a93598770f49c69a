"""GPU probe: per-sweep latency of the resident solver (difference of two max_iter-capped runs, warmed-up clocks)."""
import ctypes as C, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from colosseum_b200 import _cabi

lib = _cabi.lib()

def run(T, R, n_iter, NV=1, f64=False, pins=None, fold=0, gamma=0.99, r_const=0.0):
    S, A = T.shape[-3], T.shape[-2]
    B = 1 if T.dim() == 3 else T.shape[0]
    if NV == 4:
        B = len(pins) // 4
    vd = torch.float64 if f64 else torch.float32
    V = torch.zeros((B * NV, S), dtype=vd, device="cuda")
    status = torch.zeros(B, dtype=torch.int32, device="cuda")
    iters = torch.zeros(B * NV, dtype=torch.int64, device="cuda")
    a = _cabi.ResidentArgs()
    a.T, a.R, a.V = T.data_ptr(), (R.data_ptr() if R is not None else None), V.data_ptr()
    a.t_stride = 0 if NV == 4 else S * A * S
    a.r_stride = S * A
    a.B, a.S, a.A, a.NV, a.fold = B, S, A, NV, fold
    a.gamma, a.r_const, a.eps, a.max_iter = gamma, r_const, 0.0, n_iter
    a.pin_index = pins.data_ptr() if pins is not None else None
    a.iters_out, a.status_out = iters.data_ptr(), status.data_ptr()
    fn = lib.colo_resident_solve_f64acc if f64 else lib.colo_resident_solve_f32
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = fn(C.byref(a), torch.cuda.current_stream().cuda_stream)
    e1.record(); torch.cuda.synchronize()
    assert rc == 0, _cabi.last_error()
    return e0.elapsed_time(e1)

def main():
    # warm the clocks
    x = torch.randn(8192, 8192, device="cuda")
    for _ in range(20): y = x @ x
    torch.cuda.synchronize()
    for name in ["doc_simplegrid4", "frozenlakecontinuous_ergo0", "taxicontinuous_ergo0", "deepsea20_prand", "c2_deepsea30_prand"]:
        g = np.load(f"tests/golden/inst_{name}.npz")
        T, R = torch.from_numpy(g["T"]).cuda(), torch.from_numpy(g["R"]).cuda()
        S, A = R.shape
        cs = C.c_int(0)
        lib.colo_resident_fits(S, A, 1, 0, C.byref(cs))
        for f64 in (False, True):
            t1 = min(run(T, R, 1000, f64=f64) for _ in range(3))
            t2 = min(run(T, R, 6000, f64=f64) for _ in range(3))
            print(f"{name:28s} S={S:4d} A={A} C={cs.value:2d} NV=1 {'f64' if f64 else 'f32'}: {(t2 - t1) / 5000 * 1e3:7.3f} us/sweep  (fixed {t1 - (t2 - t1) / 5:.3f} ms)")
        K4 = (S + 3) // 4 * 4
        pins = torch.arange(K4, dtype=torch.int32, device="cuda").clamp_(max=S - 1)
        lib.colo_resident_fits(S, A, 4, 1, C.byref(cs))
        for f64 in (False, True):
            t1 = min(run(T, None, 200, NV=4, f64=f64, pins=pins, fold=2, gamma=1.0, r_const=1.0) for _ in range(2))
            t2 = min(run(T, None, 1200, NV=4, f64=f64, pins=pins, fold=2, gamma=1.0, r_const=1.0) for _ in range(2))
            print(f"{'':28s} diameter tiles={K4 // 4:4d} C={cs.value:2d} NV=4 {'f64' if f64 else 'f32'}: {(t2 - t1) / 1000 * 1e3:7.3f} us/sweep (all tiles)")


if __name__ == '__main__':
    main()
