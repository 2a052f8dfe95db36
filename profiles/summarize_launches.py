"""Per-kernel totals/shares from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections
import csv
import re
import sys

def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    tot = collections.defaultdict(float); cnt = collections.Counter()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"]).split("::")[-1]
        v = float(row["Metric Value"].replace(",", "")); unit = row["Metric Unit"]
        v = v / 1e3 if unit == "ns" else v * 1e3 if unit == "ms" else v
        tot[name] += v; cnt[name] += 1
    T = sum(tot.values())
    print(f"{'kernel':32s} {'launches':>8s} {'total ms':>10s} {'share':>7s} {'avg us':>9s}")
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        print(f"{k:32s} {cnt[k]:8d} {v / 1e3:10.3f} {v / T * 100:6.1f}% {v / cnt[k]:9.1f}")
    print(f"{'TOTAL':32s} {sum(cnt.values()):8d} {T / 1e3:10.3f}")

if __name__ == "__main__":
    main(sys.argv[1])
