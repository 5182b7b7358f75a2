"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py into profiles/ (per-kernel shares
of ONE timed step).  usage: python scripts/summarize_launches.py <launches.csv> <round-tag>"""
import collections
import csv
import re
import sys

src, tag = sys.argv[1], sys.argv[2]
lines = [l for l in open(src) if l.startswith('"')]
rows = [(x["Kernel Name"], x["Grid Size"], x["Block Size"], float(x["Metric Value"])) for x in csv.DictReader(lines)]


def short(n):
    n = re.sub(r"^void ", "", n)
    m = re.match(r"(dc::)?([A-Za-z0-9_]+)(<[^>]*>)?", n)
    return m.group(2) + (m.group(3) or "")


starts = [i for i, r in enumerate(rows) if "transpose_ncl_to_nlc" in r[0]]      # first kernel of every step
step_len = starts[1] - starts[0]
s0 = starts[3]                                                                   # a step after the warm-up passes
step = rows[s0:s0 + step_len]
tot = sum(r[3] for r in step)
agg = collections.OrderedDict()
for n, g, b, t in step:
    a = agg.setdefault(short(n), [0, 0.0])
    a[0] += 1
    a[1] += t
out = [f"# ncu launch list summary, {tag}: python bench.py --steps 2 --warmup 1 --no-cpu-baseline under ncu -c 1500 (B200, "
       "ncu --metrics gpu__time_duration.sum --clock-control none)",
       f"# the timed step = launches {s0}..{s0 + step_len - 1} of the list ({step_len} launches, {tot / 1e6:.1f} ms serialised); "
       "durations are cold-cache and serialised: compare SHARES with bench.py's `kernels`, not absolutes",
       "kernel,launches,total_ms,share"]
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append("%s,%d,%.3f,%.4f" % (k.replace(",", ";"), c, t / 1e6, t / tot))
open(f"profiles/{tag}_launches_summary.csv", "w").write("\n".join(out) + "\n")
with open(f"profiles/{tag}_launches_step.csv", "w") as f:
    f.write("idx,kernel,grid,block,duration_us\n")
    for i, (n, g, b, t) in enumerate(step):
        f.write('%d,"%s","%s","%s",%.1f\n' % (i, short(n), g, b, t / 1e3))
print("\n".join(out[:14]))
