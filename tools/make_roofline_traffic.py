#!/usr/bin/env python
"""profiles/roofline_traffic.json from an ncu launch list of one sweep step
(`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv`, tools/ncu_round2.sh):
DRAM bytes per launch of every kernel class, as bench.py's roofline.traffic reads them.
  python tools/make_roofline_traffic.py gpurun_out/r2_sweep_launches.csv > profiles/roofline_traffic.json"""
import csv
import json
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
cls = defaultdict(lambda: {"launches": 0, "time_us": 0.0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0})
inst = defaultdict(lambda: {"launches": 0, "time_us": 0.0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0})
for r in rows[1:]:
    d = dict(zip(hdr, r))
    name = d["Kernel Name"]
    base = name.replace("void ", "").split("<")[0].split("(")[0]
    if base.startswith("k_table_build"):
        base = "k_table_build"
    v = float(d["Metric Value"].replace(",", ""))
    m, u = d["Metric Name"], d["Metric Unit"]
    for tgt, key in ((cls, base), (inst, name.replace("void ", "").split("(")[0])):
        if m.startswith("gpu__time"):
            tgt[key]["time_us"] += v * {"ns": 1e-3, "us": 1, "ms": 1e3}.get(u, 1)
            tgt[key]["launches"] += 1
        elif m == "dram__bytes_read.sum":
            tgt[key]["dram_read_bytes"] += v * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        elif m == "dram__bytes_write.sum":
            tgt[key]["dram_write_bytes"] += v * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
out = {"_source": "ncu launch list of one step of the config-5 sweep (tools/ncu_round2.sh, side streams off, tables rebuilt); "
                  "cold-cache, serialised launches: compare shares, not absolutes",
       "_total_time_us": round(sum(c["time_us"] for c in cls.values()), 1)}
for k, c in cls.items():
    out[k] = {"launches_per_step": c["launches"], "time_us_per_step": round(c["time_us"], 1),
              "share_of_step": round(c["time_us"] / sum(x["time_us"] for x in cls.values()), 4),
              "dram_bytes_per_launch": int((c["dram_read_bytes"] + c["dram_write_bytes"]) / max(1, c["launches"])),
              "dram_read_bytes_per_step": int(c["dram_read_bytes"]), "dram_write_bytes_per_step": int(c["dram_write_bytes"])}
out["_by_instantiation"] = {k: {"launches": c["launches"], "time_us": round(c["time_us"], 1),
                                "dram_read_bytes": int(c["dram_read_bytes"]), "dram_write_bytes": int(c["dram_write_bytes"])}
                            for k, c in inst.items()}
print(json.dumps(out, indent=1))
