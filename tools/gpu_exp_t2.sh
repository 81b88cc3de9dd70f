#!/bin/bash
LBM_TEMPORAL=2 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -6
python - <<'PY'
import sys, os, json
sys.path.insert(0, '.')
from tools.quick_bench import run
for t2, rows in (("1", 0), ("2", 32), ("2", 64), ("2", 128)):
    os.environ["LBM_TEMPORAL"] = t2
    if rows: os.environ["LBM_T2_ROWS"] = str(rows)
    for dt, ar in (("float64", "exact"), ("float64", "fast"), ("float32", "exact")):
        r = run(4096, dt, ar, 4 if dt == "float64" else 8, steps=100)
        print(json.dumps({"temporal": t2, "t2_rows": rows, "dtype": dt, "arith": ar, "mlups": round(r[0], 1)}), flush=True)
PY
