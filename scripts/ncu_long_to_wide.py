"""ncu --csv (default page: one row per launch AND metric) -> the `--page raw` layout scripts/ncu_table.py / roofline_traffic.py
read (one row per launch, a units row under the header).  usage: ncu_long_to_wide.py in.csv out.csv"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r and "Metric Name" in r)
h = rows[hi]; c = {n: i for i, n in enumerate(h)}
launches = collections.OrderedDict(); units = {}
for r in rows[hi + 1:]:
    if len(r) < len(h): continue
    L = launches.setdefault(r[c["ID"]], {"ID": r[c["ID"]], "Kernel Name": r[c["Kernel Name"]], "Stream": r[c["Stream"]],
                                          "Block Size": r[c["Block Size"]], "Grid Size": r[c["Grid Size"]]})
    L[r[c["Metric Name"]]] = r[c["Metric Value"]]; units[r[c["Metric Name"]]] = r[c["Metric Unit"]]
metrics = list(units)
fixed = ["ID", "Kernel Name", "Stream", "Block Size", "Grid Size"]
w = csv.writer(open(sys.argv[2], "w", newline=""))
w.writerow(fixed + metrics); w.writerow([""] * len(fixed) + [units[m] for m in metrics])
for L in launches.values(): w.writerow([L.get(k, "") for k in fixed + metrics])
print(len(launches), "launches,", metrics)
