"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel total and share."""
import collections
import csv
import sys


def summarise(path, last_fraction=None):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h, data = rows[hdr], rows[hdr + 1:]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    if last_fraction:
        data = data[int(len(data) * (1 - last_fraction)):]
    agg = collections.OrderedDict()
    for r in data:
        v = float(r[vi].replace(",", ""))
        scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6}[r[ui]]
        e = agg.setdefault(r[ki][:90], [0, 0.0])
        e[0] += 1
        e[1] += v * scale
    tot = sum(v[1] for v in agg.values())
    out = [f"{'ms':>10} {'share':>6} {'count':>6}  kernel"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"{v[1]:10.3f} {100 * v[1] / tot:5.1f}% {v[0]:6d}  {k}")
    out.append(f"{tot:10.3f} total over {sum(v[0] for v in agg.values())} launches")
    return "\n".join(out)


if __name__ == "__main__":
    print(summarise(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else None))
