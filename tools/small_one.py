import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import latticeboltzmann_b200 as lb
lat = lb.Lattice(300, 200, "periodic", omega=1.0)
lat.init_equilibrium()
lat.step(50); lat.sync()
print(lat.step_timed(1000))
lat.close()
