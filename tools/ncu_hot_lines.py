"""Aggregate an `ncu --page source --csv --print-source sass,cuda` export into per-source-line shares of executed
warp instructions and stall samples.  Usage: ncu_hot_lines.py file.csv [top] [path-substring]"""
import csv
import collections
import sys


def main(path, top=25, only=None):
    rows = list(csv.reader(open(path)))
    agg = collections.OrderedDict()
    cur_file, h = None, None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            h = r
            samp, inst = h.index("# Samples"), h.index("Instructions Executed")
            continue
        if h is None or len(r) <= samp:
            continue
        try:
            ln = int(r[0])
        except ValueError:
            continue
        a = agg.setdefault((cur_file, ln), [0.0, 0.0, r[1]])
        try:
            a[0] += float(r[samp] or 0)
            a[1] += float(r[inst] or 0)
        except ValueError:
            pass
    tot = sum(a[0] for a in agg.values()) or 1.0
    toti = sum(a[1] for a in agg.values()) or 1.0
    print(f"total samples {tot:.0f}, warp instructions {toti:.0f}")
    items = [(k, a) for k, a in agg.items() if only is None or only in k[0]]
    for (f, ln), a in sorted(items, key=lambda kv: -kv[1][1])[:top]:
        print(f"{a[1] / toti * 100:5.1f}% inst {a[0] / tot * 100:5.1f}% smp  {f}:{ln:<4} {a[2].strip()[:110]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25, sys.argv[3] if len(sys.argv) > 3 else None)
