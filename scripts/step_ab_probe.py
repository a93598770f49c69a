"""A/B of the step kernels through the bench's own graph protocol (COLO_STEP_KERNEL=coop|kary)."""
import os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for kern in ("coop", "kary", "coop", "kary"):
    env = dict(os.environ, COLO_STEP_KERNEL=kern)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "step", "--steps", "400", "--cpu-seconds", "0.2"],
                         env=env, capture_output=True, text=True)
    try:
        d = json.loads(out.stdout.strip().splitlines()[-1])
        print(kern, "value %.3f G  us/step %.2f | flushed %.3f G | e2e %.3f G single %.3f G" % (
            d["value"] / 1e9, d["ms_per_step"] * 1e3, d["flushed_per_step"]["value"] / 1e9, d["e2e"]["value"] / 1e9,
            d.get("e2e_single_batch", d["e2e"])["value"] / 1e9), flush=True)
    except Exception as e:
        print(kern, "failed", e, out.stderr[-500:])
