"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel.
python scripts/launch_list_md.py launches.csv out.md "title" """
import collections
import csv
import re
import sys

src, dst, title = sys.argv[1], sys.argv[2], sys.argv[3]
agg = collections.defaultdict(lambda: [0, 0.0])
lines = [l for l in open(src) if not l.startswith("==")]
rd = csv.reader(lines)
hdr = next(rd)
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
for r in rd:
    if len(r) <= iv:
        continue
    t = float(r[iv].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[iu], 1.0)
    k = re.sub(r"\(.*", "", r[ik])[:90]
    agg[k][0] += 1
    agg[k][1] += t
tot = sum(v[1] for v in agg.values())
with open(dst, "w") as f:
    f.write(f"# {title}\n\nncu gpu__time_duration per launch, aggregated per kernel (cold-cache and serialised: compare SHARES, "
            f"not absolutes). {sum(v[0] for v in agg.values())} launches, {tot / 1e3:.1f} ms of kernel time in total.\n\n"
            "| kernel | launches | total us | mean us | share |\n|---|---|---|---|---|\n")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        f.write(f"| `{k}` | {n} | {t:.1f} | {t / n:.1f} | {100 * t / tot:.1f} % |\n")
print(open(dst).read())
