"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel; optional per-launch dump."""
import collections
import csv
import sys


def load(path):
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if r and r[0] == "ID":
            h = i
            break
    hdr = rows[h]
    ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    seq = []
    for r in rows[h + 1:]:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
        if "<" in r[ki].split("(")[0]:
            name = r[ki].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
        seq.append((name, r[gi], float(r[vi].replace(",", "")) / 1000.0))
    return seq


def main():
    seq = load(sys.argv[1])
    tot = sum(s[2] for s in seq)
    print("%d launches, %.1f us total" % (len(seq), tot))
    by = collections.defaultdict(list)
    for n, g, t in seq:
        by[n].append(t)
    print("| kernel | launches | sum us | share | min | median | max |")
    print("|---|---|---|---|---|---|---|")
    for n, v in sorted(by.items(), key=lambda kv: -sum(kv[1])):
        v2 = sorted(v)
        print("| %s | %d | %.1f | %.1f%% | %.1f | %.1f | %.1f |" % (n[:60], len(v), sum(v), 100 * sum(v) / tot, v2[0], v2[len(v2) // 2], v2[-1]))
    if len(sys.argv) > 2:
        for i, (n, g, t) in enumerate(seq):
            if sys.argv[2] in n:
                print(i, n[:50], g, "%.1f" % t)


if __name__ == "__main__":
    main()
