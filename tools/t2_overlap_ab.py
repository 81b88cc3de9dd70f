"""A/B of the overlapped pass (frame kernels on a second stream) against the serial pass, interleaved in ONE
process on two lattices of the same size so that clock / power drift hits both alike (developer tool, GPU box).

    python tools/t2_overlap_ab.py [n] [steps] [rounds]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import latticeboltzmann_b200 as lb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 6
lats = {}
for overlap in (0, 1):
    os.environ["LBM_T2_OVERLAP"] = str(overlap)
    lat = lb.Lattice(n, n, "cavity", omega=2000.0 / (0.6 * n + 1000.0), temporal=2)
    lat.init_equilibrium()
    lat.step(10)
    lat.sync()
    lats[overlap] = lat
assert lats[0].checksum() == lats[1].checksum()
ms = {0: [], 1: []}
for _ in range(rounds):
    for overlap in (0, 1):
        ms[overlap].append(lats[overlap].step_timed(steps))
assert lats[0].checksum() == lats[1].checksum()
glups = {k: [round(n * n * steps / (t * 1e-3) / 1e9, 2) for t in v] for k, v in ms.items()}
print(json.dumps({"n": n, "steps": steps, "serial": glups[0], "overlapped": glups[1],
                  "median_ratio": round(sorted(a / b for a, b in zip(ms[0], ms[1]))[rounds // 2], 4)}))
