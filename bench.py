#!/usr/bin/env python
"""Contract benchmark: D2Q9 fp64 lattice updates per second (MLUPS) of the fused time step.

    python bench.py --gpus N --steps K --warmup W [--workload NAME] [--impl reference]

N > 1 is launched by the driver as
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
(one rank per GPU; RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* from the env).

Workloads (BASELINE.json configs):
  weak16384    lid-driven cavity, 16384 x 16384 cells PER GPU, Re = 1000 (configs[4]; default:
               the configuration the 1/2/4/8-GPU metric is quoted on; N = 1 is the 1-GPU point)
  cavity4096   lid-driven cavity 4096 x 4096, Re = 1000, one GPU (configs[2])
  strong32768  lid-driven cavity 32768 x 32768 split over N GPUs (configs[3], strong scaling)

Stepping mode: `--temporal 2` (default) advances two time steps per pass over HBM (temporal blocking,
bit-identical to single steps); `--temporal 1` times the single-step kernel (one pass per step).

One JSON line on stdout (rank 0).  `value` = whole-job MLUPS with the state resident in HBM;
`e2e` = the same step driven through the host-buffer C-ABI path (pinned host f in, f out,
every step); `roofline` = the step kernel against the measured HBM peak; `cpu_baseline` =
the reference-equivalent opt2 step (numpy roll + compiled collide) on the box's host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "D2Q9 fp64 MLUPS (fused stream+collide+boundaries+halo step)"
BYTES_PER_CELL = 144          # 9 loads + 9 stores x 8 B
FALLBACK_HBM_GBS = 6650.0     # /opt/skills/guides/B200_PROFILING.md fallback


def omega_for_re(L, re=1000.0, u0=0.1):
    return 2.0 * re / (6.0 * L * u0 + re)     # slidingLid.py:28


def workload(name, n_gpus):
    """-> (global nx, ny, ndx, ndy, scaling, description)"""
    grid = {1: (1, 1), 2: (2, 1), 4: (2, 2), 8: (4, 2)}
    if n_gpus not in grid:
        raise SystemExit("--gpus must be 1, 2, 4 or 8")
    ndx, ndy = grid[n_gpus]
    if os.environ.get("LBM_BENCH_GRID"):          # developer override, e.g. "1x2"
        ndx, ndy = (int(v) for v in os.environ["LBM_BENCH_GRID"].split("x"))
    if name == "weak16384":
        n = 16384
        return n * ndx, n * ndy, ndx, ndy, "weak", "lid-driven cavity, %dx%d cells per GPU, Re=1000, %dx%d blocks" % (n, n, ndx, ndy)
    if name == "cavity4096":
        if n_gpus != 1:
            raise SystemExit("cavity4096 is a single-GPU workload")
        return 4096, 4096, 1, 1, "weak", "parallel_lid_drive_cavity 4096x4096, Re=1000"
    if name == "strong32768":
        n = 32768
        return n, n, n_gpus, 1, "strong", "lid-driven cavity 32768x32768 over %d x-slabs, Re=1000" % n_gpus
    raise SystemExit("unknown workload %r" % name)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for ts, line in self.lines:
            p = [x.strip() for x in line.split(",")]
            if len(p) < 9:
                continue
            try:
                c, m = float(p[1]), float(p[2])
            except ValueError:
                continue
            smax.append(m)
            if t0 - 0.05 <= ts <= t1 + 0.15:
                sm.append(c)
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        if not sm:   # region shorter than the sampling period: use all samples
            sm = [float(x[1].split(",")[1]) for x in self.lines if len(x[1].split(",")) >= 9] or [0.0]
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def ncu_traffic_per_launch(arith, cells):
    """dram__bytes_read.sum + dram__bytes_write.sum per step-kernel launch from the committed ncu
    captures (profiles/r01_step_kernel_dram.json: one entry per arith x cells-per-launch), or None."""
    p = os.path.join(ROOT, "profiles", "r01_step_kernel_dram.json")
    try:
        with open(p) as fh:
            for e in json.load(fh)["captures"]:
                if e["arith"] == arith and e["cells_per_launch"] == cells:
                    return e
    except Exception:
        pass
    return None


def cpu_reference(nx_total, ny_total, omega, steps, warmup, cores=None):
    """The reference-equivalent opt2 step (oracle/opt2_numpy.py) on the host cores: P processes,
    each advancing one (bx, by) block of the lattice split over P ranks (the reference's
    one-MPI-rank-per-block model, halo exchange omitted: mpirun / mpi4py are not installed).
    Returns (MLUPS, P, description)."""
    from oracle import opt2_numpy
    avail = len(os.sched_getaffinity(0))
    p = cores or max(1, min(avail, 64))
    # per-rank block of the real decomposition, bounded to 2048^2 cells so a step stays ~0.1-0.3 s
    pd = 1
    while pd * pd < p:
        pd += 1
    cap = int(os.environ.get("LBM_REF_BLOCK", "2048"))        # tests shrink the sample
    bx = int(min(cap, max(64, nx_total // pd)))
    by = int(min(cap, max(64, ny_total // pd)))
    t = opt2_numpy.run_independent_blocks(p, bx, by, omega, warmup, steps)
    mlups = p * bx * by * steps / t / 1e6
    cpu_reference.last_ms_per_step = t / steps * 1e3
    desc = "%d processes x %dx%d single-rank opt2 blocks (np.roll stream + numpy walls + compiled collide), %d steps" % (p, bx, by, steps)
    return mlups, p, desc


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    nx, ny, ndx, ndy, scaling, desc = workload(args.workload, args.gpus)
    omega = omega_for_re(nx)
    steps = max(1, min(args.steps, 10))
    mlups, cores, sample = cpu_reference(nx, ny, omega, steps, max(1, min(args.warmup, 2)))
    ms = cpu_reference.last_ms_per_step
    line = {"impl": "reference", "metric": METRIC, "value": mlups, "unit": "MLUPS", "n_gpus": args.gpus,
            "steps": steps, "warmup": max(1, min(args.warmup, 2)), "ms_per_step": ms, "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "nx": nx, "ny": ny, "omega": omega},
            "cpu_baseline": {"value": mlups, "unit": "MLUPS", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": mlups, "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def e2e_host_step(lat, steps):
    """The step through the host-buffer C-ABI path: pinned host f -> device, one fused step,
    f -> pinned host, every step (what a caller of the reference's stateless API pays).
    Rank barriers order upload / ghost refresh / step across ranks (host-side, inside the timed region)."""
    import torch
    import torch.distributed as dist
    from latticeboltzmann_b200._lib import check, c_vp
    blk, lib = lat.block, lat.block.lib
    host = torch.empty((9, blk.lnx, blk.lny), dtype=torch.float64).pin_memory()
    ptr = c_vp(host.data_ptr())
    check(lib.lb_download_f(blk.h, ptr))                       # current state as the first input

    single = dist.get_world_size() == 1

    def one():
        if single:                                             # slab-pipelined: H2D, compute, D2H overlap
            check(lib.lb_step_host(blk.h, ptr, ptr, 128))
            return
        check(lib.lb_upload_f(blk.h, ptr))                     # H2D (synchronous on return)
        dist.barrier()
        check(lib.lb_halo_refresh(blk.h))
        check(lib.lb_sync(blk.h))
        dist.barrier()
        check(lib.lb_step(blk.h, 1))
        check(lib.lb_download_f(blk.h, ptr))                   # D2H (synchronous on return)

    one()                                                      # warm-up
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dist.barrier()
    dt = time.perf_counter() - t0
    return dt / steps, host.numel() * 8


_REAL_STDOUT = None


def quiet_stdout():
    """Native libraries (NCCL's version banner, ...) print to fd 1; the contract is ONE JSON line on
    stdout.  Route fd 1 to stderr for the duration of the run and keep the real stdout for the result."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="weak16384", choices=["weak16384", "cavity4096", "strong32768"])
    ap.add_argument("--arith", default="exact", choices=["fast", "exact"],
                    help="exact: bit-identical to the reference's no-FMA build (default); fast: FMA contraction, <1e-12")
    ap.add_argument("--temporal", type=int, default=2, choices=[1, 2],
                    help="time steps per pass over HBM: 2 = temporal blocking (default, bit-identical), 1 = single-step kernel")
    ap.add_argument("--rows-per-tile", type=int, default=0)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import latticeboltzmann_b200 as lb
    from latticeboltzmann_b200 import distributed as D

    nx, ny, ndx, ndy, scaling, desc = workload(args.workload, args.gpus)
    omega = omega_for_re(nx)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d (launch with torch.distributed.run for N > 1)" % (args.gpus, world))
    rank, world, local_rank = D.init_process_group("nccl")
    numa = D.bind_to_gpu_numa(local_rank) if world > 1 else None      # pinned e2e buffers next to their GPU
    lat = D.DistributedLattice(nx, ny, ndx, ndy, "cavity", omega=omega, u_wall=0.1, dtype=np.float64,
                               arith=args.arith, device=local_rank, rows_per_tile=args.rows_per_tile or None,
                               temporal=args.temporal)
    lat.init_equilibrium()
    lat.step(args.warmup)
    lat.sync()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    launches0 = lat.block.kernel_launches
    t0 = time.time()
    ms = lat.step_timed(args.steps)         # barrier + sync, CUDA events on the launching stream, max over ranks
    t1 = time.time()
    launches = lat.block.kernel_launches - launches0
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    lat.health()

    cells = nx * ny
    mlups = cells * args.steps / (ms * 1e-3) / 1e6
    # Roofline of the dominant kernel.  One PASS over the block reads 72 B and writes 72 B per cell
    # (fp64): with the single-step kernel a pass is one time step (one launch); with temporal blocking a
    # pass is TWO time steps (the fused deep-interior kernel plus two perimeter-sized frame kernels).
    b = lat.blockinfo
    t2 = args.temporal == 2 and lat.block.temporal_active
    steps_per_pass = 2 if t2 else 1
    passes = args.steps // steps_per_pass + (args.steps % steps_per_pass)
    per_step_ms = ms / args.steps
    per_pass_ms = ms / passes
    achieved = b.lnx * b.lny * BYTES_PER_CELL / (per_pass_ms * 1e-3) / 1e9
    peak, peak_src = measured_hbm_peak()
    traffic = ncu_traffic_per_launch(args.arith + ("+t2" if t2 else ""), b.lnx * b.lny)

    e2e = None
    need = 9 * b.lnx * b.lny * 8 * world
    try:
        avail = [int(x.split()[1]) * 1024 for x in open("/proc/meminfo") if x.startswith("MemAvailable")][0]
    except Exception:
        avail = 0
    if not args.no_e2e and need * 1.5 > avail:
        e2e = {"value": None, "unit": "MLUPS", "skipped": "pinned host staging of %.0f GB exceeds 2/3 of available host memory (%.0f GB)" % (need / 1e9, avail / 1e9)}
    elif not args.no_e2e:
        try:
            sec, nbytes = e2e_host_step(lat, args.e2e_steps)
            sec = D.max_over_ranks(sec)
            e2e = {"value": cells / sec / 1e6, "unit": "MLUPS", "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes,
                   "steps": args.e2e_steps, "path": ("lb_step_host: pinned host f[9,nx,ny] in/out every step, 128 slabs, H2D/compute/D2H overlapped" if world == 1 else
                            "lb_upload_f + lb_halo_refresh + lb_step(1) + lb_download_f on pinned host f[9,lnx,lny] per rank; bytes are per rank")}
        except Exception as exc:      # reported, never hidden
            e2e = {"value": None, "unit": "MLUPS", "error": repr(exc)}

    cpu = None
    if rank == 0 and args.gpus == 1 and not args.no_cpu_baseline:
        from oracle import opt2_numpy
        v, cores, sample = cpu_reference(nx, ny, omega, 3, 1)
        cap = int(os.environ.get("LBM_REF_BLOCK", "2048"))
        one = opt2_numpy.run_independent_blocks(1, cap, cap, omega, 1, 2)
        cpu = {"value": v, "unit": "MLUPS", "cores": cores, "kind": "port", "sample": sample,
               "one_core_value": cap * cap * 2 / one / 1e6,
               "note": "value: reference-structured opt2 step (np.roll stream + numpy walls + compiled collide) on all cores used; "
                       "one_core_value: the same on 1 core"}

    lat.close()
    if rank == 0:
        line = {"metric": METRIC, "value": mlups, "unit": "MLUPS", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": per_step_ms, "higher_is_better": True, "scaling": scaling,
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": desc, "nx": nx, "ny": ny, "ndx": ndx, "ndy": ndy, "omega": omega, "u0": 0.1,
                           "arith": args.arith, "steps_per_hbm_pass": steps_per_pass,
                           "single_step_roofline_mlups": peak * 1e9 / BYTES_PER_CELL / 1e6,
                           "halo": "in-kernel peer stores over NVLink (CUDA IPC), device-side flags",
                           "numa_bound_cpus": (len(numa) if numa else None),
                           "l2": "inputs larger than L2 (%.1f GB per buffer per GPU, A/B ping-pong)" % (9 * b.lnx * b.lny * 8 / 1e9)},
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches * world,
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": (traffic or {}).get("dram_bytes_per_launch"), "peak_source": peak_src,
                             "bytes_per_cell_per_pass": BYTES_PER_CELL, "steps_per_pass": steps_per_pass,
                             "cells_per_launch": b.lnx * b.lny, "ms_per_pass": per_pass_ms,
                             "kernel": "t2_interior_kernel (+ 2 frame kernels per pass)" if t2 else "step_kernel",
                             "traffic_note": (traffic or {}).get("note")},
                "cpu_baseline": cpu}
        emit(line)
    import torch.distributed as dist
    if dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
