import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import latticeboltzmann_b200 as lb
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
lat = lb.Lattice(n, n, "cavity", omega=1.0, arith=sys.argv[2] if len(sys.argv) > 2 else "exact")
lat.blocks[0].set_temporal(2, 64)
lat.init_equilibrium()
lat.step(8); lat.sync()
print(n * n * 20 / lat.step_timed(20) / 1e3, "MLUPS")
lat.close()
