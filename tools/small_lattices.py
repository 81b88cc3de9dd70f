"""Timing of the L2-resident BASELINE configurations (configs[0]: 300 x 200 shear wave with the per-step
amplitude probe; configs[1]: 512 x 512 simple_flows) on the resident multi-step kernel and on the per-step path.
Developer tool (GPU box)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import latticeboltzmann_b200 as lb


def timed(lat, n):
    lat.step(64)
    lat.sync()
    best = 1e30
    for _ in range(3):
        t0 = time.perf_counter()
        lat.step(n)
        lat.sync()
        best = min(best, time.perf_counter() - t0)
    lat.health()
    return best / n


def case(name, nx, ny, boundary, n, probe=False, **kw):
    out = {"case": name, "cells": nx * ny}
    for mode in (1, 0):
        os.environ["LBM_RESIDENT"] = str(mode)
        lat = lb.Lattice(nx, ny, boundary, **kw)
        lat.init_equilibrium()
        if probe:
            lat.probe_shear_enable(0.01 * np.sin(2 * np.pi * np.arange(nx) / nx), 4 * n + 64)
        sec = timed(lat, n)
        lat.close()
        key = "resident" if mode else "per_step"
        out[key + "_us_per_step"] = round(sec * 1e6, 3)
        out[key + "_glups"] = round(nx * ny / sec / 1e9, 2)
    print(json.dumps(out), flush=True)


case("shear_wave_300x200_probe", 300, 200, "periodic", 4000, probe=True, omega=1.0)
case("shear_wave_300x200", 300, 200, "periodic", 4000, omega=1.0)
case("shear_wave_300x300_fp32", 300, 300, "periodic", 4000, omega=0.3, dtype=np.float32)
case("couette_512x514", 512, 514, "sf_couette", 2000, omega=0.5, u_wall=0.1)
case("poiseuille_514x514", 514, 514, "sf_poiseuille", 2000, omega=0.5, u_wall=0.0, rho_in=1.001, rho_out=0.999)
case("sliding_lid_514x514", 514, 514, "sf_sliding_lid", 2000, omega=1.0, u_wall=0.1)
case("cavity_512x512", 512, 512, "cavity", 2000, omega=1.7)
case("cavity_1024x1024", 1024, 1024, "cavity", 1000, omega=1.7)
