"""Launch-bound regime: steps/s of the small BASELINE configs."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import latticeboltzmann_b200 as lb

def t(nx, ny, boundary, n=2000, **kw):
    lat = lb.Lattice(nx, ny, boundary, omega=1.0, **kw)
    lat.init_equilibrium()
    lat.step(200); lat.sync()
    ms = lat.step_timed(n)
    lat.close()
    print("%-16s %5dx%-5d %8.2f us/step %10.1f MLUPS" % (boundary, nx, ny, ms * 1e3 / n, nx * ny * n / ms / 1e3), flush=True)

t(300, 200, "periodic")
t(300, 300, "periodic")
t(512, 512, "cavity")
t(512, 514, "sf_couette", u_wall=0.1)
t(514, 514, "sf_poiseuille", u_wall=0.0, rho_in=1.001, rho_out=0.999)
t(514, 514, "sf_sliding_lid", u_wall=0.1)
t(1024, 1024, "cavity")
t(2048, 2048, "cavity")
