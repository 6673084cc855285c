"""Condense an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel name, launches and time."""
import csv
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = OrderedDict()
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] in ("ns", "nsecond") else v * 1e3 if r[ui] in ("ms", "msecond") else v
    name = r[ki][:90]
    a = agg.setdefault(name, [0, 0.0, 0.0])
    a[0] += 1
    a[1] += v
    a[2] = v
print("%-90s %6s %12s %10s" % ("kernel", "calls", "total us", "last us"))
for name, (n, tot, last) in agg.items():
    print("%-90s %6d %12.1f %10.1f" % (name, n, tot, last))
