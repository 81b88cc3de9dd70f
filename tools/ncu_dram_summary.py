"""Turn `ncu --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum` logs of short
`bench.py` runs into profiles/r02_kernel_dram.json (the `roofline.traffic` source of bench.py).

usage: python tools/ncu_dram_summary.py out.json arith:cells:log.csv [...]

`arith` ending in "+t2" is the two-steps-per-pass mode: ONE PASS = t2_frame1 + t2_interior + t2_frame2 (the log must
hold whole passes); otherwise one launch of step_kernel is one pass.
"""
import csv
import json
import sys

SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6}


def launches(path, needle):
    """[(kernel name, grid, read bytes, written bytes, ns)] of the launches whose name contains `needle`, in order."""
    by = {}
    for r in csv.reader(open(path)):
        if len(r) > 10 and needle in r[4] and r[0].isdigit():
            d = by.setdefault(int(r[0]), {"name": r[4], "grid": r[8]})
            d[r[-3]] = float(r[-1].replace(",", "")) * SCALE.get(r[-2], 1)
    return [(d["name"], d["grid"], d["dram__bytes_read.sum"], d["dram__bytes_write.sum"], d["gpu__time_duration.sum"])
            for _id, d in sorted(by.items())]


def mean(rows, j):
    return sum(r[j] for r in rows) / len(rows)


caps = []
for spec in sys.argv[2:]:
    arith, cells, path = spec.split(":")
    cap = {"arith": arith, "cells_per_launch": int(cells), "algorithmic_bytes_per_launch": int(cells) * 144}
    if arith.endswith("+t2"):
        parts = {k: launches(path, k) for k in ("t2_frame1", "t2_interior", "t2_frame2")}
        n = min(len(v) for v in parts.values())
        assert n > 0, "no complete pass in " + path
        parts = {k: v[-n:] for k, v in parts.items()}
        rd = sum(mean(v, 2) for v in parts.values())
        wr = sum(mean(v, 3) for v in parts.values())
        ns = sum(mean(v, 4) for v in parts.values())
        it = parts["t2_interior"]
        cap.update(launches_averaged=n, dram_bytes_read=rd, dram_bytes_write=wr, dram_bytes_per_launch=rd + wr, gpu_time_ns=ns,
                   interior_kernel={"grid": it[0][1], "dram_bytes_read": mean(it, 2), "dram_bytes_write": mean(it, 3),
                                    "gpu_time_ns": mean(it, 4)},
                   frame_kernels_ns=mean(parts["t2_frame1"], 4) + mean(parts["t2_frame2"], 4),
                   note="ONE PASS = two time steps = t2_frame1 + t2_interior (grid %s) + t2_frame2, bytes and times summed; under ncu the "
                        "kernels run one at a time (in production the two frame kernels run concurrently with the interior kernel); "
                        "ncu --clock-control none on a short `bench.py` run; read excess over the algorithmic bytes = the one-cell "
                        "level-(n+1) halo of a fused tile (2 rows per tile, 2 columns per 254); source %s" % (it[0][1], path))
    else:
        rows = launches(path, "step_kernel")
        rd, wr, ns = mean(rows, 2), mean(rows, 3), mean(rows, 4)
        cap.update(launches_averaged=len(rows), dram_bytes_read=rd, dram_bytes_write=wr, dram_bytes_per_launch=rd + wr, gpu_time_ns=ns,
                   note="single-step kernel, one launch per step; ncu --clock-control none on a short `bench.py --temporal 1` run; source " + path)
    caps.append(cap)
json.dump({"captures": caps}, open(sys.argv[1], "w"), indent=1)
print(json.dumps(caps, indent=1))
