"""Aggregates `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` by CUDA source line: samples, warp instructions and
the leading stall reasons of the hottest lines.   python tools/ncu_lines.py file.csv [top]"""
import csv
import sys
from collections import defaultdict


def main():
    path, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25
    rows = list(csv.reader(open(path)))
    cur_file, hdr = None, None
    agg = defaultdict(lambda: defaultdict(float))
    src = {}
    for r in rows:
        if len(r) == 2 and r[0] in ("File Name", "File Path"):
            cur_file = r[1].split("/")[-1]
            continue
        if r and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) != len(hdr) or not r[0].isdigit():
            continue
        key = (cur_file, int(r[0]))
        if r[2] == "":  # the CUDA line itself
            src[key] = r[1].strip()[:110]
            continue
        for h, v in zip(hdr[4:], r[4:]):
            try:
                agg[key][h] += float(v)
            except ValueError:
                pass
    tot_s = sum(a["# Samples"] for a in agg.values())
    tot_i = sum(a["Instructions Executed"] for a in agg.values())
    print(f"total samples {tot_s:.0f}, warp instructions {tot_i:.0f}")
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["# Samples"])[:top]:
        lead = sorted(((a[s], s[6:]) for s in stalls), reverse=True)[:3]
        print(f"{key[0]}:{key[1]:4d} samp {100 * a['# Samples'] / max(tot_s, 1):5.1f}% inst {100 * a['Instructions Executed'] / max(tot_i, 1):5.1f}%  "
              + " ".join(f"{n}={v:.0f}" for v, n in lead) + "  | " + src.get(key, ""))


if __name__ == "__main__":
    main()
