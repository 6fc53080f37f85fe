"""Prints (duration us, kernel, grid, block) of an ncu `--metrics gpu__time_duration.sum --csv` launch list.
Usage: launch_table.py file.csv [first [last]]"""
import csv, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
seq = []
for x in csv.DictReader(lines):
    if x.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(x["Metric Value"].replace(",", "")); u = x["Metric Unit"]
    v = v / 1e3 if u in ("nsecond", "ns") else v * 1e3 if u in ("msecond", "ms") else v
    seq.append((x["Kernel Name"][:64], v, x.get("Grid Size", ""), x.get("Block Size", "")))
a = int(sys.argv[2]) if len(sys.argv) > 2 else 0
b = int(sys.argv[3]) if len(sys.argv) > 3 else len(seq)
for k, v, g, bl in seq[a:b]:
    print(f"{v:9.1f} us  {k} {g} {bl}")
