"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list: python tools/launch_summary.py <csv> <out.txt> [title]"""
import csv
import re
import sys
from collections import defaultdict

src, out = sys.argv[1], sys.argv[2]
title = sys.argv[3] if len(sys.argv) > 3 else src
rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
hdr = rows[0]
kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[kn].replace("(anonymous namespace)::", "").replace("<unnamed>::", ""))
    name = re.sub(r"<.*", "", name)
    v = float(r[mv].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[mu], 1e-3)
    tot[name] += v
    cnt[name] += 1
allt = sum(tot.values())
with open(out, "w") as f:
    f.write(f"# {title}\n# per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes\n")
    f.write(f"{'kernel':60s} {'launches':>8s} {'total_us':>14s} {'share':>7s} {'avg_us':>12s}\n")
    for k in sorted(tot, key=tot.get, reverse=True):
        f.write(f"{k:60s} {cnt[k]:8d} {tot[k]:14.1f} {100 * tot[k] / allt:6.2f}% {tot[k] / cnt[k]:12.1f}\n")
print(open(out).read())
