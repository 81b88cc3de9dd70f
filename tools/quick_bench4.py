import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.quick_bench import run
tag = os.environ.get("LBM_NATIVE_LIB", "product").split("/")[-1]
for dtype in ("float32", "float64"):
    for arith in ("exact", "fast"):
        for rows in (4, 8, 16):
            m, gbs = run(4096, dtype, arith, rows)
            print(json.dumps({"lib": tag, "dtype": dtype, "arith": arith, "rows": rows, "mlups": round(m, 1), "GBs": round(gbs, 1)}), flush=True)
