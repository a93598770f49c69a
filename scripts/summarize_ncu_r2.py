"""Turns the round-2 ncu captures (gpurun_out/r2_*.ncu-rep, scripts/profile_r2.sh) into the tracked summaries:
profiles/r2_<name>_metrics.csv (the judged metrics of every captured launch), profiles/r2_ncu_summary.json (what bench.py
quotes), profiles/ncu_traffic.json (DRAM bytes per unit for roofline.traffic), profiles/r2_launches.md (the launch
list of the default bench command, aggregated per kernel) and profiles/r2_sass_<kernel>.txt (cuobjdump -sass of the
shipped library for the hot kernels, with the mnemonic histogram on top).  Runs here, no GPU needed."""
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "sm__cycles_active.avg", "sm__cycles_elapsed.avg",
]
UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "Tbyte": 1e12}


def raw_rows(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def main():
    global PROF
    if len(sys.argv) > 1:  # on the GPU box: write next to the reports (only gpurun_out/ travels back)
        PROF = os.path.join(ROOT, sys.argv[1])
    os.makedirs(PROF, exist_ok=True)
    summary, traffic = {}, {}
    for rep in sorted(f for f in os.listdir(OUT) if f.startswith("r2_") and f.endswith(".ncu-rep")):
        name = rep[3:-8]
        hdr, units, data = raw_rows(os.path.join(OUT, rep))
        cols = [w for w in WANT if w in hdr]
        with open(os.path.join(PROF, f"r2_{name}_metrics.csv"), "w", newline="") as f:
            wr = csv.writer(f)
            wr.writerow(["kernel"] + cols)
            wr.writerow(["(unit)"] + [units[hdr.index(c)] for c in cols])
            for r in data:
                wr.writerow([r[hdr.index("Kernel Name")][:90]] + [r[hdr.index(c)] for c in cols])
        r = data[-1]

        def val(k, scale_bytes=False):
            if k not in hdr:
                return None
            x = float(r[hdr.index(k)].replace(",", ""))
            return x * UNIT.get(units[hdr.index(k)], 1.0) if scale_bytes else x

        dur = val("gpu__time_duration.sum")
        du = units[hdr.index("gpu__time_duration.sum")]
        dur_us = dur * {"us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}.get(du, 1.0)
        rd, wrb = val("dram__bytes_read.sum", True), val("dram__bytes_write.sum", True)
        summary[name] = {
            "kernel": r[hdr.index("Kernel Name")][:80], "duration_us": dur_us,
            "registers_per_thread": val("launch__registers_per_thread"), "grid": val("launch__grid_size"),
            "waves_per_sm": val("launch__waves_per_multiprocessor"),
            "dram_bytes": (rd or 0) + (wrb or 0), "dram_gbs": ((rd or 0) + (wrb or 0)) / dur_us / 1e3,
            "dram_pct_of_peak": val("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            "lts_pct_of_peak": val("lts__throughput.avg.pct_of_peak_sustained_elapsed"),
            "l2_hit_pct": val("lts__t_sector_hit_rate.pct"), "l1_hit_pct": val("l1tex__t_sector_hit_rate.pct"),
            "issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "warps_active_pct": val("sm__warps_active.avg.pct_of_peak_sustained_active"),
            "warp_instructions": val("smsp__inst_executed.sum"),
            "tensor_pipe_active_pct": val("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
            "stall_long_scoreboard": val("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
        }
        if name.startswith("step_"):
            n_env = int(name.split("_")[1])
            summary[name]["warp_instructions_per_env_step"] = summary[name]["warp_instructions"] * 32 / n_env / 32 * 32 / 32
            summary[name]["warp_instructions_per_env_step"] = summary[name]["warp_instructions"] / (n_env / 32.0)
            summary[name]["dram_bytes_per_env_step"] = summary[name]["dram_bytes"] / n_env
            if n_env == 65536:
                traffic["step"] = {"dram_bytes_per_unit": summary[name]["dram_bytes"] / n_env, "unit": "env-step",
                                   "from": f"profiles/r2_{name}_metrics.csv"}
        if name == "backup_c4":
            traffic["backup_c4"] = {"dram_bytes_per_unit": summary[name]["dram_bytes"] / 4096, "unit": "MDP-sweep (S=512,A=4)",
                                    "from": f"profiles/r2_{name}_metrics.csv"}
        if name == "backup_c5":
            traffic["backup_c5"] = {"dram_bytes_per_unit": summary[name]["dram_bytes"] / (16384.0 * 8 * 16384), "unit": "T element (captured at S=16,384)",
                                    "from": f"profiles/r2_{name}_metrics.csv"}
    json.dump(summary, open(os.path.join(PROF, "r2_ncu_summary.json"), "w"), indent=1)
    if traffic:
        json.dump(traffic, open(os.path.join(PROF, "ncu_traffic.json"), "w"), indent=1)
    # launch list of the default bench command
    ll = os.path.join(OUT, "r2_launches.csv")
    if os.path.isfile(ll):
        agg = collections.defaultdict(lambda: [0, 0.0])
        lines = [l for l in open(ll) if not l.startswith("==")]
        rd = csv.reader(lines)
        hdr = next(rd)
        ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
        for r in rd:
            if len(r) <= iv:
                continue
            t = float(r[iv].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[iu], 1.0)
            k = re.sub(r"\(.*", "", r[ik])[:70]
            agg[k][0] += 1
            agg[k][1] += t
        tot = sum(v[1] for v in agg.values())
        with open(os.path.join(PROF, "r2_launches.md"), "w") as f:
            f.write("# Launch list of `python bench.py --steps 20 --warmup 5 --no-sweep` (ncu gpu__time_duration, first 3000 "
                    "launches; cold-cache and serialised: compare SHARES, not absolutes)\n\n| kernel | launches | total us | share |\n|---|---|---|---|\n")
            for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
                f.write(f"| `{k}` | {n} | {t:.1f} | {100 * t / tot:.1f} % |\n")
    # SASS of the hot kernels
    so = os.path.join(ROOT, "colosseum_b200", "_lib", "libcolosseum_b200.so")
    syms = subprocess.run(["cuobjdump", "-elf", so], capture_output=True, text=True).stdout
    for tag, pat in (("backup_kernel_f32_max_vec_warp", r"_ZN4colo13backup_kernelIfLi0ELb1ELb0EEEv\w+"),
                     ("gs_solve_tma_kernel_f32", r"_ZN4colo19gs_solve_tma_kernelIfEEv\w+"),
                     ("env_step_kary_lean_kernel_4", r"_ZN4colo25env_step_kary_lean_kernelILi4EEEv\w+"),
                     ("hitting_umma_kernel_128", r"_ZN4colo19hitting_umma_kernelILi128EEEv\w+")):
        m = re.search(pat, syms)
        if not m:
            print("symbol not found:", tag)
            continue
        sass = subprocess.run(["cuobjdump", "-sass", "-fun", m.group(0), so], capture_output=True, text=True).stdout
        body = [l for l in sass.splitlines() if re.match(r"\s+/\*[0-9a-f]{4}\*/", l)]
        hist = collections.Counter()
        for l in body:
            t = re.sub(r"^\s*/\*[0-9a-f]+\*/\s*", "", l).split()
            op = t[1] if t[0].startswith("@") else t[0]
            hist[op.split(".")[0].rstrip(";")] += 1
        with open(os.path.join(PROF, f"r2_sass_{tag}.txt"), "w") as f:
            f.write(f"# cuobjdump -sass -fun {m.group(0)} libcolosseum_b200.so  ({len(body)} instructions)\n# mnemonic histogram: "
                    + ", ".join(f"{k} {v}" for k, v in hist.most_common(24)) + "\n")
            f.write("\n".join(re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", l) for l in body) + "\n")
    print(json.dumps(summary, indent=1)[:3000])


if __name__ == "__main__":
    main()
