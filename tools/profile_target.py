"""ncu target: a few passes of the stepping kernels on one cavity lattice (not a benchmark).
usage: profile_target.py n temporal steps [dtype]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import latticeboltzmann_b200 as lb

n, temporal, steps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
dtype = np.dtype(sys.argv[4]) if len(sys.argv) > 4 else np.float64
lat = lb.Lattice(n, n, "cavity", omega=2000.0 / (0.6 * n + 1000.0), dtype=dtype, temporal=temporal)
lat.blocks[0].set_use_graph(False)
lat.init_equilibrium()
lat.step(steps)
lat.sync()
lat.health()
lat.close()
