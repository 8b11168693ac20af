"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals/shares."""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[hdr]
    ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v *= {"us": 1e-3, "ns": 1e-6, "s": 1e3, "ms": 1.0}.get(r[ui], 1.0)
        agg.setdefault(r[ki].split("(")[0][:70], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    print(f"# {path}: {sum(len(v) for v in agg.values())} launches, {tot:.3f} ms total (cold-cache, serialised: compare shares)")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k:72s} n={len(v):4d} total={sum(v):10.3f} ms avg={sum(v) / len(v):9.4f} ms share={sum(v) / tot * 100:5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
