"""ncu launch list (gpu__time_duration.sum per launch, CSV) -> markdown summary grouped by kernel.
usage: python scripts/launches_md.py gpurun_out/r2_launches.csv profiles/r2_launches.md "<command line that was profiled>" """
import collections
import csv
import re
import shutil
import sys

src, dst, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
rows = []
with open(src) as f:
    lines = [ln for ln in f if not ln.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        rows.append((r["Kernel Name"], float(r["Metric Value"]) / 1e6))  # ns -> ms
agg = collections.OrderedDict()
for name, ms in rows:
    short = re.sub(r"^void ", "", name)
    short = re.sub(r"\(.*$", "", short)
    a = agg.setdefault(short, [0, 0.0])
    a[0] += 1
    a[1] += ms
ours = {k: v for k, v in agg.items() if k.startswith("lob::")}
other = {k: v for k, v in agg.items() if not k.startswith("lob::")}
tot = sum(v[1] for v in ours.values())
out = [f"# ncu launch list: `{cmd}`", "",
       f"{len(rows)} launches; per-launch times under ncu are serialised and cold-cache: use the SHARES, not the absolutes.",
       f"Raw list: `{dst.replace('.md', '.csv').split('/')[-1]}`.", "",
       "| kernel | launches | total ms | share of liblob time |", "|---|---:|---:|---:|"]
for k, (n, ms) in sorted(ours.items(), key=lambda kv: -kv[1][1]):
    out.append(f"| `{k}` | {n} | {ms:.2f} | {100 * ms / tot:.1f} % |")
out += ["", "Other (not liblob; synthetic-input generation and torch glue): " +
        ", ".join(f"`{k[:60]}` {v[1]:.1f} ms" for k, v in sorted(other.items(), key=lambda kv: -kv[1][1])[:6])]
open(dst, "w").write("\n".join(out) + "\n")
shutil.copy(src, dst.replace(".md", ".csv"))
print("wrote", dst, f"liblob total {tot:.1f} ms")
