"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (share of device time)."""
import collections
import csv
import re
import sys


def main(path, top=30):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    n = 0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)
        key = re.sub(r"\(.*", "", re.sub(r"<.*", "", row["Kernel Name"])).replace("void ", "").strip()
        tot[key] += v
        cnt[key] += 1
        n += 1
    T = sum(tot.values())
    print(f"# {path}: {n} launches, {T/1e3:.2f} ms of kernel time (cold-cache, serialised: compare shares)")
    print(f"{'us':>12s} {'share':>7s} {'n':>5s}  kernel")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:top]:
        print(f"{v:12.1f} {100*v/T:6.1f}% {cnt[k]:5d}  {k}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
