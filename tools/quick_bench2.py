import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.quick_bench import run
import numpy as np
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
for dtype in ("float64", "float32"):
    for arith in ("exact", "fast"):
        for rows in (4, 8, 16):
            m, gbs = run(n, dtype, arith, rows)
            print(json.dumps({"n": n, "dtype": dtype, "arith": arith, "rows": rows, "mlups": round(m, 1), "GBs": round(gbs, 1)}), flush=True)
