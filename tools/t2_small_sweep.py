"""Where does two-steps-per-pass start to win?  Small lattices: fused tiles of 8/16/32 rows vs single-step launches (graph replay)
vs the resident kernel (developer tool, GPU box)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import latticeboltzmann_b200 as lb


def rate(n, steps, **kw):
    lat = lb.Lattice(n, n, "cavity", omega=1.7, **kw)
    lat.init_equilibrium()
    lat.step(256)
    lat.sync()
    best = min(lat.step_timed(steps) for _ in range(3))
    lat.health()
    lat.close()
    return round(n * n * steps / (best * 1e-3) / 1e9, 2)


for n in (512, 768, 1024, 1536, 2048, 2560):
    res = {"n": n}
    steps = 1024
    for rows in (8, 16, 32):
        os.environ["LBM_T2_ROWS"] = str(rows)
        os.environ["LBM_RESIDENT"] = "0"
        res["t2_rows%d" % rows] = rate(n, steps, temporal=2)
    os.environ.pop("LBM_T2_ROWS")
    res["single_graph"] = rate(n, steps, temporal=1)
    os.environ["LBM_RESIDENT"] = "1"
    if n * n <= 1 << 20:
        res["resident"] = rate(n, steps, temporal=1)
    print(json.dumps(res), flush=True)
