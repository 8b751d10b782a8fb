"""Aggregate an `ncu --page source --csv` export (SASS+source view) into per-source-line stall-sample shares."""
import csv
import collections
import sys


def main(path, top=25):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
    h = rows[hi]
    samp = h.index("# Samples")
    inst = h.index("Instructions Executed")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= samp:
            continue
        try:
            ln = int(r[0])
        except ValueError:
            continue
        a = agg.setdefault(ln, [0.0, 0.0, r[1]])
        try:
            a[0] += float(r[samp] or 0)
            a[1] += float(r[inst] or 0)
        except ValueError:
            pass
    tot = sum(a[0] for a in agg.values()) or 1.0
    toti = sum(a[1] for a in agg.values()) or 1.0
    print(f"total samples {tot:.0f}, warp instructions {toti:.0f}")
    for ln, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{a[0] / tot * 100:5.1f}% smp {a[1] / toti * 100:5.1f}% inst  L{ln:<4} {a[2].strip()[:120]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
