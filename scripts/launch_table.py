"""Print an ncu launch list (csv of `--metrics gpu__time_duration.sum`) as kernel / grid / block / microseconds."""
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1000 if u in ("nsecond", "ns") else (v * 1000 if u in ("msecond", "ms") else v)
    print("%-72s %-16s %-12s %10.1f us" % (row["Kernel Name"][:72], row["Grid Size"], row["Block Size"], v))
