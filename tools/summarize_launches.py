"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ik])
    name = re.sub(r"^void |t2p::|\(anonymous namespace\)::", "", name)
    v = float(r[iv].replace(",", ""))
    u = r[iu]
    ms = v / 1e6 if u in ("ns", "nsecond") else (v / 1e3 if u in ("us", "usecond") else v)
    agg[name][0] += 1
    agg[name][1] += ms
tot = sum(v[1] for v in agg.values())
print(f"{'kernel':70s} {'n':>5s} {'ms':>9s} {'share':>7s}")
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:70]:70s} {n:5d} {ms:9.3f} {100 * ms / tot:6.1f}%")
print(f"{'TOTAL':70s} {sum(v[0] for v in agg.values()):5d} {tot:9.3f}")
