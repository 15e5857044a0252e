"""Per-kernel totals of an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file L ...`).
    python tools/launch_summary.py L.csv > summary.csv
Columns: kernel, launches, total_us, mean_us, share (of the captured device time).
"""
import csv
import sys
from collections import OrderedDict


def main():
    rows = []
    with open(sys.argv[1], newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    kn, mn, mv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    tot = OrderedDict()
    for r in rd:
        if r[mn] != "gpu__time_duration.sum":
            continue
        name = r[kn].split("(")[0]
        t = tot.setdefault(name, [0, 0.0])
        t[0] += 1
        t[1] += float(r[mv].replace(",", "")) / 1000.0
    all_us = sum(t[1] for t in tot.values())
    w = csv.writer(sys.stdout)
    w.writerow(["kernel", "launches", "total_us", "mean_us", "share"])
    for name, (c, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        w.writerow([name, c, "%.1f" % us, "%.2f" % (us / c), "%.4f" % (us / all_us)])


if __name__ == "__main__":
    main()
