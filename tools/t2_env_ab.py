"""Interleaved A/B/... of library settings on lattices of one size, in ONE process, so that clock / power drift hits
all variants alike (developer tool, GPU box).  Each variant is a comma-separated list of environment settings read
by lb_create (LBM_T2_OVERLAP, LBM_T2_ROWS, ...); "-" is the default configuration.

    python tools/t2_env_ab.py n steps rounds VARIANT [VARIANT ...]
    python tools/t2_env_ab.py 16384 100 6 LBM_T2_OVERLAP=0 -          # serial pass vs overlapped frames
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import latticeboltzmann_b200 as lb

n, steps, rounds = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
variants = sys.argv[4:]
lats = []
for v in variants:
    env = dict(kv.split("=") for kv in v.split(",") if "=" in kv)
    os.environ.update(env)
    lat = lb.Lattice(n, n, "cavity", omega=2000.0 / (0.6 * n + 1000.0), temporal=int(env.get("TEMPORAL", 2)))
    for k in env:
        os.environ.pop(k)
    lat.init_equilibrium()
    lat.step(10)
    lat.sync()
    lats.append(lat)
assert len({lat.checksum() for lat in lats}) == 1
ms = [[] for _ in lats]
for _ in range(rounds):
    for i, lat in enumerate(lats):
        ms[i].append(lat.step_timed(steps))
assert len({lat.checksum() for lat in lats}) == 1
for v, t in zip(variants, ms):
    g = sorted(n * n * steps / (x * 1e-3) / 1e9 for x in t)
    print(json.dumps({"n": n, "steps": steps, "variant": v, "glups_median": round(g[len(g) // 2], 2),
                      "glups": [round(x, 2) for x in g]}), flush=True)
