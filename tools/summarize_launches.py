#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (count, total, share)."""
import csv
import re
import sys
from collections import defaultdict

path, out = sys.argv[1], sys.argv[2]
rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
hdr = rows[0]
ik, iv, ig, ib = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
agg = defaultdict(lambda: [0, 0.0])
order = []
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ik]).replace("void ", "").strip()
    name = re.sub(r"<unnamed>::", "", name)
    ns = float(r[iv].replace(",", ""))
    agg[name][0] += 1
    agg[name][1] += ns
    order.append((name, r[ig], r[ib], ns))
tot = sum(v[1] for v in agg.values())
with open(out, "w") as f:
    f.write("# ncu launch list summary (%s)\n\n" % path.split("/")[-1])
    f.write("Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n\n")
    f.write("| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write("| `%s` | %d | %.3f | %.1f%% |\n" % (k[:90], n, t / 1e6, 100 * t / tot))
    f.write("\nTotal %.3f ms over %d launches.\n\n## Largest launches\n\n| kernel | grid | block | ms |\n|---|---|---|---:|\n" % (tot / 1e6, len(order)))
    for name, g, b, ns in sorted(order, key=lambda x: -x[3])[:12]:
        f.write("| `%s` | %s | %s | %.3f |\n" % (name[:70], g, b, ns / 1e6))
print(open(out).read())
