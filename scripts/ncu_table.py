"""ncu csv (raw page, one row per launch) -> per-kernel table: launches, total / mean time, DRAM bytes, achieved DRAM GB/s.
usage: ncu_table.py launches.csv [peak_gbs]"""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
peak = float(sys.argv[2]) if len(sys.argv) > 2 else 6551.7
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hi]
units = rows[hi + 1]
col = {n: i for i, n in enumerate(h)}
def val(r, name):
    v = float(r[col[name]].replace(",", "")); u = units[col[name]]
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "": 1.0}.get(u, 1.0)
    return v * scale
agg = collections.OrderedDict()
for r in rows[hi + 2:]:
    if len(r) < len(h): continue
    name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void ", "").replace("scn::", "")
    a = agg.setdefault(name, [0, 0.0, 0.0, 0.0, 0.0, 0.0])
    t = val(r, "gpu__time_duration.sum")
    a[0] += 1; a[1] += t
    a[2] += val(r, "dram__bytes_read.sum"); a[3] += val(r, "dram__bytes_write.sum")
    if "lts__t_bytes.sum" in col: a[4] += val(r, "lts__t_bytes.sum")
    a[5] = max(a[5], t)
tot = sum(a[1] for a in agg.values())
print("| kernel | launches | total us | share | mean us | max us | DRAM MB (r+w) | DRAM GB/s | frac of %.0f | L2 MB |" % peak)
print("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    gbs = (a[2] + a[3]) / (a[1] * 1e-6) / 1e9 if a[1] else 0
    print("| `%s` | %d | %.1f | %.1f %% | %.1f | %.1f | %.1f | %.0f | %.3f | %.1f |" % (
        name[:70], a[0], a[1], 100 * a[1] / tot, a[1] / a[0], a[5], (a[2] + a[3]) / 1e6, gbs, gbs / peak, a[4] / 1e6))
print("\ntotal %.1f us over %d launches" % (tot, sum(a[0] for a in agg.values())))
