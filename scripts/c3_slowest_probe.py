"""Per-phase wall times (COLO_SUITE_VERBOSE) of the most expensive C3 instances, one at a time on one worker.
python scripts/c3_slowest_probe.py [n_slowest]"""
import os
import sys

os.environ["COLO_SUITE_VERBOSE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from colosseum_b200.suite import load_suite_all, run_many_native, suite_size

g = os.path.join(ROOT, "tests", "golden")
n = min(suite_size(g), 128)
suite = load_suite_all(g, indices=list(range(n)))


def cost(inst):
    m = len(inst.nodes) if inst.episodic and inst.nodes is not None else inst.S
    return float(m) * m * inst.tables.A


order = sorted(suite, key=lambda i: -cost(i))[: int(sys.argv[1]) if len(sys.argv) > 1 else 10]
run_many_native([(order[0], 0)], n_workers=1, n_envs=1024, n_steps=10)
for inst in order:
    print(f"--- {inst.name} S={inst.S} A={inst.A} H={inst.H} nodes={0 if inst.nodes is None else len(inst.nodes)}", file=sys.stderr, flush=True)
    run_many_native([(inst, 0)], n_workers=1, n_envs=1024, n_steps=1000)
