"""Per-kernel share of device time from an ncu launch list
   (ncu --metrics gpu__time_duration.sum --clock-control none -c N --csv --log-file launches.csv <command>).
usage: python profiles/launch_shares.py launches.csv > shares.txt"""
import csv, sys
from collections import defaultdict
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows[1:]:
    v = float(r[iv].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iu], 1e-6)
    name = r[ik].split("(")[0][:90]
    tot[name] += v; cnt[name] += 1
allms = sum(tot.values())
print("%-92s %8s %12s %8s %10s" % ("kernel", "launches", "total_ms", "share", "us/launch"))
for k in sorted(tot, key=lambda k: -tot[k]):
    print("%-92s %8d %12.3f %7.2f%% %10.1f" % (k, cnt[k], tot[k], 100 * tot[k] / allms, 1e3 * tot[k] / cnt[k]))
