"""ncu raw CSV (ncu -i report.ncu-rep --page raw --csv) -> per-kernel summary JSON: average duration, DRAM bytes,
DRAM / L2 throughput, tensor-pipe activity.   python tools/ncu_summary.py raw.csv out.json"""
import collections
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}


def num(r, name):
    i = col.get(name)
    if i is None or r[i] in ("", "n/a"):
        return None
    try:
        return float(r[i].replace(",", ""))
    except ValueError:
        return None


units = rows[1]
acc = collections.OrderedDict()
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    d = acc.setdefault(name, collections.defaultdict(list))
    for key, metric in (("us", "gpu__time_duration.sum"), ("dram_read", "dram__bytes_read.sum"),
                        ("dram_write", "dram__bytes_write.sum"),
                        ("dram_pct", "dram__throughput.avg.pct_of_peak_sustained_elapsed"),
                        ("l2_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
                        ("tensor_pct_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                        ("sm_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed")):
        v = num(r, metric)
        if v is not None:
            scale = 1.0
            u = units[col[metric]] if metric in col else ""
            if key == "us":
                scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1.0)
            if key.startswith("dram_r") or key.startswith("dram_w"):
                scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
            d[key].append(v * scale)
out = {}
for name, d in acc.items():
    rec = {"launches": len(d["us"])}
    for k, v in d.items():
        rec[k] = round(sum(v) / len(v), 3)
    if "dram_read" in rec:
        rec["dram_bytes"] = round(rec["dram_read"] + rec.get("dram_write", 0.0))
    out[name[:120]] = rec
json.dump(out, open(sys.argv[2], "w"), indent=1)
for k, v in out.items():
    print("%-90s %s" % (k[:90], {a: b for a, b in v.items() if a in ("launches", "us", "dram_bytes", "dram_pct", "tensor_pct_active")}))
