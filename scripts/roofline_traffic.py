"""profiles/roofline_traffic.json from an ncu per-launch csv of one training step (scripts/ncu_step.py train):
DRAM bytes (read + write) per launch of the three kernels bench.py reports a roofline for, level-0 launches only.
usage: roofline_traffic.py gpurun_out/r2c_train.csv profiles/<name of the committed copy>"""
import csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hi]
col = {n: i for i, n in enumerate(h)}


def sel(name, grid=None, min_us=0.0):
    out = []
    for r in rows[hi + 2:]:
        if len(r) < len(h) or name not in r[col["Kernel Name"]]:
            continue
        if grid is not None and not r[col["Grid Size"]].replace(" ", "").startswith(grid):
            continue
        if float(r[col["gpu__time_duration.sum"]]) / 1e3 < min_us:
            continue
        out.append((float(r[col["gpu__time_duration.sum"]]) / 1e3,
                    float(r[col["dram__bytes_read.sum"]]) + float(r[col["dram__bytes_write.sum"]])))
    return out


def entry(name, grid, what, min_us=0.0, largest=False):
    x = sel(name, grid, min_us)
    if largest:      # the level-0 launch = the one that moves the most bytes
        x = [max(x, key=lambda t: t[1])]
    return {"kernel": what, "launches": len(x), "bytes_per_launch": sum(b for _, b in x) / len(x),
            "us_per_launch_under_ncu": sum(t for t, _ in x) / len(x)}


src = sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]
out = {"source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum "
                 "--clock-control none, one training step of the bench scene (scripts/ncu_step.py train): " + src,
       "conv": entry("k_conv_ts<32", None, "k_conv_ts<32,1,4>, level-0 SubM 3^3 32->32 launches (forward and input gradient)"),
       "wgrad": entry("k_wgrad_ts<32>", None, "k_wgrad_ts<32>, level-0 32->32 launches (+ k_wgrad_ts_reduce<32>: see the csv)"),
       "rulebook": entry("k_subm_map", None, "k_subm_map, level 0 (the largest of the six launches)", largest=True)}
json.dump(out, open("profiles/roofline_traffic.json", "w"), indent=1)
print(json.dumps(out, indent=1))
