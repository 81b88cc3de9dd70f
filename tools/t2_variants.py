"""Developer tool: build A/B variants of the fused two-step kernel next to the product library
(latticeboltzmann_b200/csrc/variants/lib_<name>.so) and, on a GPU box, check + time each of them.

    python tools/t2_variants.py build            (here, no GPU)
    python tools/t2_variants.py run [n] [steps]  (GPU box; one subprocess per variant: LBM_NATIVE_LIB)
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VDIR = os.path.join(ROOT, "latticeboltzmann_b200", "csrc", "variants")

VARIANTS = {
    "f32_3x2": ["LBM_T2_STAGES_F32=3", "LBM_T2_MINB_F32=2"],
    "f32_6x2": ["LBM_T2_STAGES_F32=6", "LBM_T2_MINB_F32=2"],
    "f32_6x3": ["LBM_T2_STAGES_F32=6", "LBM_T2_MINB_F32=3"],
    "f32_4x4": ["LBM_T2_STAGES_F32=4", "LBM_T2_MINB_F32=4"],
    "f32_8x2": ["LBM_T2_STAGES_F32=8", "LBM_T2_MINB_F32=2"],
    "res256": ["LBM_RES_THREADS=256"],
    "res512": ["LBM_RES_THREADS=512"],
    "res1024": ["LBM_RES_THREADS=1024"],
    "tma1": ["LBM_T2_TMA=1", "LBM_T2_STAGES=1", "LBM_T2_MINB=4"],
    "tma2": ["LBM_T2_TMA=1", "LBM_T2_STAGES=2", "LBM_T2_MINB=3"],
    "tma4": ["LBM_T2_TMA=1", "LBM_T2_STAGES=4", "LBM_T2_MINB=2"],
    "tma3": ["LBM_T2_TMA=1", "LBM_T2_STAGES=3", "LBM_T2_MINB=2"],
    "r1": ["LBM_T2_TMA=0", "LBM_T2_COMPACT_RING=0", "LBM_T2_MINB=4"],                # round-1 shipped kernel (direct loads)
    "compact": ["LBM_T2_TMA=0", "LBM_T2_COMPACT_RING=1", "LBM_T2_MINB=4"],
    "r1_al": ["LBM_T2_TMA=0", "LBM_T2_COMPACT_RING=0", "LBM_T2_MINB=4", "LBM_T2_W=252", "LBM_T2_S=0", "LBM_T2_OFF=2"],
}


def build(names):
    from latticeboltzmann_b200.build import build_native
    os.makedirs(VDIR, exist_ok=True)
    for name in names:
        out = os.path.join(VDIR, "lib_%s.so" % name)
        build_native(defines=VARIANTS[name], out=out)
        print("built", out, flush=True)


CHILD = r'''
import json, os, sys
sys.path.insert(0, %(root)r)
import numpy as np
import latticeboltzmann_b200 as lb
from oracle import oracle as orc
name, n, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
res = {"variant": name, "n": n}
# bit-exactness on ragged multi-tile shapes (forced temporal blocking), fp64 and fp32
ok = True
for (nx, ny, dt, bc) in ((70, 530, np.float64, "cavity"), (131, 1031, np.float64, "cavity"), (45, 777, np.float32, "periodic")):
    f0 = orc.perturbed_state(nx, ny, dt, seed=3)
    ref = f0.copy()
    (orc.cavity_run if bc == "cavity" else orc.periodic_run)(ref, 1.7, 9)
    lat = lb.Lattice(nx, ny, bc, omega=1.7, dtype=dt, temporal=2)
    lat.upload(f0); lat.step(9); got = lat.download(); lat.health(); lat.close()
    ok = ok and bool(np.array_equal(got, ref))
res["bit_exact"] = ok
for rows in (32, 64):
    os.environ["LBM_T2_ROWS"] = str(rows)
    lat = lb.Lattice(n, n, "cavity", omega=2000.0 / (0.6 * n + 1000.0), temporal=2,
                     dtype=np.float32 if os.environ.get("LBM_VAR_DTYPE") == "f32" else np.float64)
    lat.init_equilibrium(); lat.step(6); lat.sync()
    best = min(lat.step_timed(steps) for _ in range(3))
    lat.health(); lat.close()
    res["glups_rows%%d" %% rows] = round(n * n * steps / (best * 1e-3) / 1e9, 2)
print(json.dumps(res), flush=True)
'''


def run(n, steps, names):
    for name in names:
        so = os.path.join(VDIR, "lib_%s.so" % name)
        env = dict(os.environ, LBM_NATIVE_LIB=so)
        r = subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT}, name, str(n), str(steps)], env=env,
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
        print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else "(no output) rc=%d" % r.returncode, flush=True)


if __name__ == "__main__":
    cmd = sys.argv[1] if len(sys.argv) > 1 else "build"
    names = [a for a in sys.argv[2:] if a in VARIANTS] or list(VARIANTS)
    nums = [int(a) for a in sys.argv[2:] if a.isdigit()]
    if cmd == "build":
        build(names)
    else:
        run(nums[0] if nums else 16384, nums[1] if len(nums) > 1 else 20, names)
