"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections
import csv
import sys


def main(path, steps=2):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    tot = 0.0
    for row in csv.DictReader(lines):
        name = row["Kernel Name"][:64]
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)
        d = agg.setdefault(name, [0, 0.0])
        d[0] += 1
        d[1] += v
        tot += v
    print("%-66s %5s %11s %6s" % ("kernel", "n", "us/step", "share"))
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-66s %5d %11.1f %5.1f%%" % (k, n // steps, v / steps, 100 * v / tot))
    print("total us/step %.1f" % (tot / steps))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 2)
