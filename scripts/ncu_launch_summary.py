"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list; with --seq also
the launches of the last full step in order.   python scripts/ncu_launch_summary.py list.csv [--seq N]"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = None
d = collections.defaultdict(list); seq = []
for r in rows:
    if "Kernel Name" in r: hdr = r; continue
    if hdr and len(r) == len(hdr):
        rr = dict(zip(hdr, r))
        try:
            us = float(rr["Metric Value"]) / 1000
        except ValueError:
            continue
        name = rr["Kernel Name"].split("(")[0][-28:]
        d[name].append(us); seq.append((name, us, rr["Stream"], rr["Grid Size"]))
for k, v in d.items(): print("%-30s n=%3d avg %7.1f min %7.1f max %7.1f us" % (k, len(v), sum(v) / len(v), min(v), max(v)))
if "--seq" in sys.argv:
    n = int(sys.argv[sys.argv.index("--seq") + 1])
    for name, us, st, grid in seq[-n:]: print("  %-28s %7.1f us  stream %s grid %s" % (name, us, st, grid))
