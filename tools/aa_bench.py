"""In-place (AA pattern) vs A/B single-step kernel: rate at 16384^2 and footprint at 32768^2 (developer tool, GPU box)."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import latticeboltzmann_b200 as lb

out = {}
for name, kw in (("ab_single_step", dict(temporal=1)), ("inplace_aa", dict(inplace=True))):
    n = 16384
    lat = lb.Lattice(n, n, "cavity", omega=2000.0 / (0.6 * n + 1000.0), **kw)
    lat.init_equilibrium()
    lat.step(6)
    lat.sync()
    ms = min(lat.step_timed(20) for _ in range(3))
    d = lat.checksum()
    lat.health()
    lat.close()
    out[name] = {"glups_16384": round(n * n * 20 / ms / 1e6, 2), "gbs": round(n * n * 144 * 20 / ms / 1e6, 1), "digest_after_66": "%016x" % d}
free0 = torch.cuda.mem_get_info(0)[0]
n = 32768
lat = lb.Lattice(n, n, "cavity", omega=2000.0 / (0.6 * n + 1000.0), inplace=True)
used = free0 - torch.cuda.mem_get_info(0)[0]
lat.init_equilibrium()
lat.step(4)
lat.sync()
ms = lat.step_timed(10)
lat.health()
lat.close()
out["inplace_aa_32768"] = {"device_bytes": used, "gb": round(used / 1e9, 2), "glups": round(n * n * 10 / ms / 1e6, 2)}
print(json.dumps(out))
