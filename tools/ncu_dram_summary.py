"""Turn `ncu --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum` logs
into profiles/r01_step_kernel_dram.json (the `roofline.traffic` source of bench.py).
usage: python tools/ncu_dram_summary.py out.json arith:cells:log.csv [...]"""
import csv
import json
import sys


def parse(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and "step_kernel" in r[4]]
    by = {}
    for r in rows:
        by.setdefault(r[0], {})[r[-3]] = (float(r[-1].replace(",", "")), r[-2])
    out = []
    for _id, m in by.items():
        def val(name):
            v, unit = m[name]
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6}.get(unit, 1)
            return v * scale
        out.append((val("dram__bytes_read.sum"), val("dram__bytes_write.sum"), val("gpu__time_duration.sum")))
    return out


caps = []
for spec in sys.argv[2:]:
    arith, cells, path = spec.split(":")
    launches = parse(path)
    rd = sum(x[0] for x in launches) / len(launches)
    wr = sum(x[1] for x in launches) / len(launches)
    ns = sum(x[2] for x in launches) / len(launches)
    caps.append({"arith": arith, "cells_per_launch": int(cells), "launches_averaged": len(launches),
                 "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
                 "algorithmic_bytes_per_launch": int(cells) * 144, "gpu_time_ns": ns,
                 "note": "ncu --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum per launch, source " + path})
json.dump({"captures": caps}, open(sys.argv[1], "w"), indent=1)
print(json.dumps(caps, indent=1))
