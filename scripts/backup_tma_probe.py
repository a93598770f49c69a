"""A/B of the synchronous sweep on C4: LDG streaming kernel vs the TMA-staged variant (COLO_BACKUP_TMA=1), one process per
configuration (the switches are read once).  python scripts/backup_tma_probe.py [B]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    sys.path.insert(0, ROOT)
    import torch

    import bench
    import colosseum_b200.dynamic_programming as dp

    B = int(sys.argv[2])
    T, R = bench.make_c4_batch(B, 512, 4, seed=100)
    vi = dp.BatchedValueIteration(T, R, gamma=0.99, precision="f32")
    vi.sweep(5)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    vi.sweep(50)
    e1.record()
    e1.synchronize()
    ms = e0.elapsed_time(e1) / 50
    bytes_ = B * (4 * 512 * 4 * 512 + 4 * 512 * 4 + 8 * 512 + 4 * 512 * 4)
    print(f"{ms:.4f} ms/sweep {bytes_ / ms / 1e6:.0f} GB/s V checksum {float(vi.values.double().sum()):.6f}")
    sys.exit(0)
B = sys.argv[1] if len(sys.argv) > 1 else "4096"
for tag, env in (("ldg (shipped)", {}), ("tma 24 stages, 1 CTA/SM", {"COLO_BACKUP_TMA": "1"}),
                 ("tma 12 stages, 2 CTA/SM", {"COLO_BACKUP_TMA": "1", "COLO_BACKUP_TMA_CTAS": "2"})):
    try:
        out = subprocess.run([sys.executable, __file__, "--child", B], env={**os.environ, **env}, capture_output=True,
                             text=True, timeout=150)
        msg = out.stdout.strip() or ([l for l in out.stderr.splitlines() if "rror" in l] or ["?"])[-1][:300]
    except subprocess.TimeoutExpired:
        msg = "timed out after 150 s"
    print(f"{tag:44s}: {msg}", flush=True)
