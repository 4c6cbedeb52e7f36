"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv`) of bench.py: per-kernel launches, time and share over the
COMPLETE steps found in the capture (a step starts at the patch-unfold kernel).  Usage: python tools/ncu_launch_summary.py launches.csv"""
import collections, csv, re, sys

rows = []
with open(sys.argv[1]) as f:
    lines = [ln for ln in f if not ln.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v
    rows.append((r["Kernel Name"], us))
starts = [i for i, (n, _) in enumerate(rows) if "im2col_patch" in n]
if len(starts) < 2:
    sys.exit(f"need at least two step starts in the capture, found {len(starts)} in {len(rows)} launches")
sel = rows[starts[0]:starts[-1]]
steps = len(starts) - 1
tot = sum(us for _, us in sel)
print(f"{len(rows)} launches captured; {steps} complete step(s) = {len(sel)} launches ({len(sel) / steps:.0f} per step), "
      f"{tot / 1e3 / steps:.2f} ms per step under ncu (cold-cache, serialised per-launch times: compare SHARES)")
agg = collections.OrderedDict()
for n, us in sel:
    n = re.sub(r"^void ", "", n)
    n = re.sub(r"\(.*$", "", n).replace("wg::<unnamed>::", "").replace("(anonymous namespace)::", "")
    n = re.sub(r"\(int\)", "", n)
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += us
for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n[:72]:72s} launches/step={c / steps:6.1f} ms/step={us / 1e3 / steps:8.3f} share={100 * us / tot:5.1f}%")
