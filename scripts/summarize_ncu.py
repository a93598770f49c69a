"""Turns the ncu artefacts brought back in gpurun_out/ into the tracked summaries under profiles/.

    python scripts/summarize_ncu.py <round-tag> <launch-list.csv> <name=report.ncu-rep@units_per_launch> [...]

(units per launch = envs for the step kernel, MDP instances for the backup kernel, in the profiled command)

Writes profiles/<tag>_launches.md (per-kernel launch counts / durations / share of the step),
profiles/<tag>_<name>_metrics.csv (the raw-page metrics that matter) and updates profiles/ncu_traffic.json
(dram read+write bytes per launch, which bench.py reports as roofline.traffic)."""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROF = os.path.join(ROOT, "profiles")
KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__waves_per_multiprocessor", "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def launches(tag, path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except Exception:
            continue
        agg.setdefault(row["Kernel Name"], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    ours = sum(sum(v) for k, v in agg.items() if "colo::" in k)
    out = [f"# {tag}: ncu launch list (`--metrics gpu__time_duration.sum --clock-control none`)", "",
           "Cold-cache, serialised launches: compare SHARES, not absolutes.  Source: " + os.path.basename(path), "",
           "| kernel | launches | avg us | total ms | share of all | share of colo:: |", "|---|---:|---:|---:|---:|---:|"]
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        name = k.replace("|", "/")[:110]
        share_o = f"{100 * sum(v) / ours:.1f}%" if "colo::" in k and ours else "-"
        out.append(f"| `{name}` | {len(v)} | {sum(v) / len(v) / 1e3:.2f} | {sum(v) / 1e6:.3f} | {100 * sum(v) / tot:.1f}% | {share_o} |")
    open(os.path.join(PROF, f"{tag}_launches.md"), "w").write("\n".join(out) + "\n")


def report(tag, name, rep, n_units):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    traffic_path = os.path.join(PROF, "ncu_traffic.json")
    traffic = json.load(open(traffic_path)) if os.path.isfile(traffic_path) else {}
    with open(os.path.join(PROF, f"{tag}_{name}_metrics.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "metric", "value", "unit"])
        for r in rows[2:]:
            kname = r[hdr.index("Kernel Name")]
            for k in KEEP:
                if k in hdr:
                    w.writerow([kname[:100], k, r[hdr.index(k)], units[hdr.index(k)]])
            for i, k in enumerate(hdr):
                if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and r[i]:
                    try:
                        if float(r[i]) >= 0.1:
                            w.writerow([kname[:100], k, r[i], units[i]])
                    except ValueError:
                        pass
            def val(k):
                v, u = float(r[hdr.index(k)].replace(",", "")), units[hdr.index(k)].lower()
                return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
            key = kname.split("<")[0].split("::")[-1].replace("void ", "").strip()
            tot = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
            traffic[key] = {"dram_bytes_per_launch": tot, "units_per_launch": n_units, "dram_bytes_per_unit": tot / n_units,
                            "source": f"profiles/{tag}_{name}_metrics.csv"}
            traffic[name] = dict(traffic[key], kernel=kname[:100])  # also under the spec's own name (e.g. backup_c5)
    json.dump(traffic, open(traffic_path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    os.makedirs(PROF, exist_ok=True)
    tag = sys.argv[1]
    launches(tag, sys.argv[2])
    for spec in sys.argv[3:]:
        name, rep = spec.split("=")
        rep, units = rep.split("@")
        report(tag, name, rep, float(units))
    print(open(os.path.join(PROF, "ncu_traffic.json")).read())
