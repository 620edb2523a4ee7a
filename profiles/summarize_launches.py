"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: python profiles/summarize_launches.py gpurun_out/launches.csv [skip_first_n]"""
import csv
import re
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ns = v * {"ns": 1, "us": 1e3, "usecond": 1e3, "nsecond": 1, "ms": 1e6, "msecond": 1e6}.get(unit, 1)
        rows.append((int(r["ID"]), r["Kernel Name"], ns))
    rows = [r for r in rows if r[0] >= skip]
    agg = defaultdict(lambda: [0.0, 0])
    for _, name, ns in rows:
        short = re.sub(r"\(.*", "", name)
        short = re.sub(r"^void ", "", short)
        short = short[:70]
        agg[short][0] += ns
        agg[short][1] += 1
    total = sum(v[0] for v in agg.values())
    print(f"{len(rows)} launches, {total / 1e3:.1f} us total (cold-cache, serialised: compare shares)")
    print(f"{'kernel':72s} {'launches':>8s} {'total us':>10s} {'avg us':>8s} {'share':>7s}")
    for k, (ns, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"{k:72s} {n:8d} {ns / 1e3:10.1f} {ns / n / 1e3:8.2f} {100 * ns / total:6.1f}%")


if __name__ == "__main__":
    main()
