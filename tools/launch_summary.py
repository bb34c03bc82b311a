"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, mean us, share.
   python tools/launch_summary.py <launches.csv> [first_fraction_to_skip]"""
import collections
import csv
import io
import sys


def main(path, skip=0.0):
    rows = [l for l in open(path) if l.startswith('"')]
    r = list(csv.DictReader(io.StringIO("".join(rows))))
    r = r[int(len(r) * skip):]
    agg = collections.OrderedDict()
    for x in r:
        k = x["Kernel Name"].split("(")[0][-64:]
        v = float(x["Metric Value"].replace(",", ""))
        u = x["Metric Unit"]
        v = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
        agg.setdefault(k, [0, 0.0])
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    print("%d launches, %.1f us total" % (len(r), tot))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-66s %4d %9.1f %6.3f" % (k, v[0], v[1] / v[0], v[1] / tot))


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 0.0)
