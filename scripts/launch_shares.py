"""Kernel shares of one step from an `ncu --metrics gpu__time_duration.sum --csv` launch list.

    python scripts/launch_shares.py profiles/r02_launches_bench_resnet50.csv [launches_per_step]

The list holds every launch of the run (compile, warm-up, timed steps, per-op profile passes); the LAST complete forward is
located by its ingest kernel and summarised by kernel function: launches, total us, share of the step."""
import collections
import csv
import re
import sys

path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    lines = [ln for ln in f if not ln.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        rows.append((re.sub(r"\(.*", "", r["Kernel Name"]), us))
starts = [i for i, (n, _) in enumerate(rows) if "ingest" in n or "nchw_to_nhwc" in n]
# a forward = from one ingest launch to the next; take the most common length among the eager steps
lens = collections.Counter(b - a for a, b in zip(starts, starts[1:]))
L = lens.most_common(1)[0][0]
cands = [a for a, b in zip(starts, starts[1:]) if b - a == L]
a = cands[-1]
step = rows[a:a + L]
tot = sum(t for _, t in step)
agg = collections.OrderedDict()
for n, t in step:
    agg.setdefault(n, [0, 0.0])
    agg[n][0] += 1
    agg[n][1] += t
print(f"# {path}: last complete forward = launches {a}..{a + L - 1} ({L} launches, {tot:.1f} us under ncu: cold caches, serialised)")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n[:100]:100s} x{c:3d} {t:9.1f} us  {100 * t / tot:5.1f} %")
