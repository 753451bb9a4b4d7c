"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name (used by tools/gpu_run.sh)."""
import csv
import re
import sys
from collections import defaultdict

rows = list(csv.reader(l for l in open(sys.argv[1], errors="replace") if l.startswith('"')))
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
ui = hdr.index("Metric Unit")
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows[1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
    name = re.sub(r"\(.*", "", r[ki])
    tot[name] += v
    cnt[name] += 1
all_ms = sum(tot.values())
print("total %.3f ms in %d launches" % (all_ms, sum(cnt.values())))
for k in sorted(tot, key=lambda k: -tot[k]):
    print("%10.3f ms %5.1f%% %6d x  %s" % (tot[k], 100 * tot[k] / all_ms, cnt[k], k[:110]))
