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
    captures (profiles/r02_kernel_dram.json: one entry per arith x cells-per-launch), or None."""
    p = os.path.join(ROOT, "profiles", "r02_kernel_dram.json")
    try:
        with open(p) as fh:
            for e in json.load(fh)["captures"]:
                if e["arith"] == arith and e["cells_per_launch"] == cells:
                    return e
    except Exception:
        pass
    return None


def cpu_reference(nx_total, ny_total, omega, steps, warmup, cores=None):
    """The reference's parallel cavity as it runs under mpirun (cavity_opt2.py:214-277), on the host cores:
    P processes in a Cartesian grid, one block each with the reference's ghost layers, and per step
    communicate() -> stream_and_bounce_back() (np.roll + numpy walls) -> compiled collide
    (oracle/opt2_numpy.run_decomposed; mpirun / mpi4py are not installed, so the Sendrecv of communicate() are
    shared-memory copies ordered by process barriers).  The lattice is the workload's own when it has at most
    16384^2 cells, else a 16384^2 sample of it (bounded so that a step stays around a second).
    Returns (MLUPS, P, description)."""
    from oracle import opt2_numpy
    avail = len(os.sched_getaffinity(0))
    p = 1
    while p * 2 <= (cores or max(1, min(avail, 64))):
        p *= 2
    pdx = 1
    while pdx * pdx * 2 <= p:
        pdx *= 2
    pdy = p // pdx
    cap = int(os.environ.get("LBM_REF_BLOCK", "4096"))        # per-process block edge; tests shrink it
    nx = int(min(nx_total, 16384, cap * pdx))
    ny = int(min(ny_total, 16384, cap * pdy))
    t, _ = opt2_numpy.run_decomposed(pdx, pdy, nx, ny, omega, warmup, steps)
    mlups = nx * ny * steps / t / 1e6
    cpu_reference.last_ms_per_step = t / steps * 1e3
    whole = nx == nx_total and ny == ny_total
    desc = ("%s%dx%d lattice split over %dx%d processes (one block each, ghost layers and the four-way ghost exchange of "
            "communicate() every step through shared memory, np.roll stream + numpy walls + compiled collide), %d steps"
            % ("the whole " if whole else "a %d-cell sample of the %d-cell workload: " % (nx * ny, nx_total * ny_total),
               nx, ny, pdx, pdy, steps))
    return mlups, p, desc


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    nx, ny, ndx, ndy, scaling, desc = workload(args.workload, args.gpus)
    omega = omega_for_re(nx)
    steps = max(1, min(args.steps, 50))       # ~0.3 s per step on 16 cores: the requested count whenever it fits a minute
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
    """The step through the host-buffer C-ABI path: pinned host f -> device, one fused step, f -> pinned host,
    every step (what a caller of the reference's stateless API pays).  One block: lb_step_host, slab-pipelined.
    Decomposed: DistributedLattice.step_host = rim upload, rank barrier, rim pushed into the neighbours' ghosts
    over NVLink, rank barrier, then every rank's slab pipeline -- the barriers are inside the timed region."""
    import torch
    import torch.distributed as dist
    from latticeboltzmann_b200._lib import check, c_vp
    blk, lib = lat.block, lat.block.lib
    host_t = torch.empty((9, blk.lnx, blk.lny), dtype=torch.float64).pin_memory()
    host = host_t.numpy()
    check(lib.lb_download_f(blk.h, c_vp(host_t.data_ptr())))           # current state as the first input

    lat.step_host(host, host, 128)                                      # warm-up (creates streams / events / staging)
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        lat.step_host(host, host, 128)
    dist.barrier()
    dt = time.perf_counter() - t0
    return dt / steps, host_t.numel() * 8


def parity_selfcheck(D, ndx, ndy, device):
    """Before anything is timed: the deterministic ragged cases of latticeboltzmann_b200/selfcheck.py through the
    SAME decomposition and halo path as the benchmark (forced temporal blocking, odd step count => double steps +
    one single step), gathered and compared by SHA-256 with the CPU oracle's result committed under
    tests/golden/bench_parity.json (tests/make_bench_parity.py).  The oracle itself is not imported here."""
    import torch.distributed as dist
    from latticeboltzmann_b200 import selfcheck
    with open(os.path.join(ROOT, "tests", "golden", "bench_parity.json")) as fh:
        golden = json.load(fh)["sha256"]
    out = {"bit_exact": True, "ndx": ndx, "ndy": ndy, "cases": {},
           "golden": "tests/golden/bench_parity.json (CPU oracle, tests/make_bench_parity.py)"}
    for name, (boundary, nx, ny, dtype, omega, u0, steps) in selfcheck.CASES.items():
        lat = D.DistributedLattice(nx, ny, ndx, ndy, boundary, omega=omega, u_wall=u0, dtype=np.dtype(dtype),
                                   arith="exact", device=device, temporal=2)
        b = lat.blockinfo
        lat.init_equilibrium(*selfcheck.fields(nx, ny, dtype, b.x0, b.y0, b.lnx, b.lny))
        t2 = lat.block.temporal_active
        lat.step(steps)
        g = lat.gather_f(0)
        lat.health()
        lat.close()
        if dist.get_rank() == 0:
            ok = selfcheck.digest(g) == golden[name]
            out["cases"][name] = {"bit_exact": ok, "temporal_blocking": bool(t2), "steps": steps}
            out["bit_exact"] = out["bit_exact"] and ok
    return out


def timed_lattice(D, nx, ny, ndx, ndy, omega, dtype, arith, device, temporal, warmup, steps, rows_per_tile=None):
    """init -> warmup -> `steps` timed steps (barrier + sync, CUDA events, max over ranks) -> digest.  Returns
    (lattice still open, ms, global digest after warmup + steps)."""
    lat = D.DistributedLattice(nx, ny, ndx, ndy, "cavity", omega=omega, u_wall=0.1, dtype=dtype, arith=arith,
                               device=device, rows_per_tile=rows_per_tile, temporal=temporal)
    lat.init_equilibrium()
    lat.step(warmup)
    lat.sync()
    ms = lat.step_timed(steps)
    lat.health()
    return lat, ms, lat.checksum()


def strip_rows(lnx, tile_rows, rows=64):
    """A strip of `rows` rows in the middle of the block centred on a row seam of the fused tiles (tiles start at
    row 2 + tile_rows * m; with tiles of at most 32 rows the strip straddles two seams)."""
    a = 2 + tile_rows * max(1, (lnx // 2) // tile_rows) - rows // 2
    return (a, a + rows) if a - 2 >= 0 and a + rows + 2 <= lnx else None


def strip_check_vs_oracle(pre, post, omega, u0):
    """cpu_baseline leg only (the one place bench.py may run oracle/): two oracle steps on the downloaded strip
    `pre` (rows a-2 .. a+rows+2, all columns, level n) must reproduce `post` (rows a .. a+rows, level n+2)
    bit for bit.  The strip is interior in x, so the oracle runs it without left/right walls; its wrapped edge
    rows are garbage after each step and are not compared."""
    from oracle import oracle as orc
    t1, t2 = np.empty_like(pre), np.empty_like(pre)
    orc.cavity_step_pull(pre, t1, omega, u0=u0, walls_lr=False)
    orc.cavity_step_pull(t1, t2, omega, u0=u0, walls_lr=False)
    return bool(np.array_equal(t2[:, 2:-2], post))


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
    ap.add_argument("--no-extras", action="store_true", help="skip the parity self-check, the single-step comparison and the extra workloads (profiling runs)")
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

    # ---- parity first: ragged lattices through this decomposition, against the committed oracle hashes --------
    parity = None if args.no_extras else parity_selfcheck(D, ndx, ndy, local_rank)

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
    digest_main = None if args.no_extras else lat.checksum()     # global state after warmup + steps (device-side digest)

    cells = nx * ny
    mlups = cells * args.steps / (ms * 1e-3) / 1e6
    # Roofline of the dominant kernel.  One PASS over the block reads 72 B and writes 72 B per cell
    # (fp64): with the single-step kernel a pass is one time step (one launch); with temporal blocking a
    # pass is TWO time steps (the fused deep-interior kernel plus two perimeter-sized frame kernels).
    b = lat.blockinfo
    t2 = args.temporal == 2 and lat.block.temporal_active
    steps_per_pass = 2 if t2 else 1
    tile_rows = lat.block.temporal_rows if t2 else None
    passes = args.steps // steps_per_pass + (args.steps % steps_per_pass)
    per_step_ms = ms / args.steps
    per_pass_ms = ms / passes
    achieved = b.lnx * b.lny * BYTES_PER_CELL / (per_pass_ms * 1e-3) / 1e9
    peak, peak_src = measured_hbm_peak()
    traffic = ncu_traffic_per_launch(args.arith + ("+t2" if t2 else ""), b.lnx * b.lny)

    # Strip of the benchmarked state for the oracle (checked in the cpu_baseline leg below): level n rows with a
    # two-row halo, one more pass of the stepping kernel, level n+2 rows.
    strip = None
    if rank == 0 and args.gpus == 1 and not args.no_cpu_baseline and not args.no_extras:
        rows = strip_rows(b.lnx, lat.block.temporal_rows)
        if rows:
            pre = lat.block.download_rows(rows[0] - 2, rows[1] + 2)
            lat.step(2)
            lat.sync()
            strip = (rows, pre, lat.block.download_rows(rows[0], rows[1]))

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
                            "per rank: lb_step_host_begin (rim up) | barrier | lb_halo_refresh over NVLink | barrier | lb_step_host (128 slabs, H2D/compute/D2H "
                            "overlapped) on pinned host f[9,lnx,lny]; bytes are per rank")}
        except Exception as exc:      # reported, never hidden
            e2e = {"value": None, "unit": "MLUPS", "error": repr(exc)}
    lat.close()

    # ---- the stated metric: the SINGLE-STEP kernel (144 B per cell per step) on the same workload, same step
    # count; its final state must have the same digest as the temporal-blocking run (two steps per pass ==
    # single steps, bit for bit, at the benchmarked size).
    single = None
    extra = {}
    free_b = torch.cuda.mem_get_info(local_rank)[0]
    lattice_b = 2 * 9 * (b.lnx + 2) * (b.lny + 64) * 8
    if not args.no_extras and t2:
        if lattice_b * 1.02 < free_b:
            lat1, ms1, digest1 = timed_lattice(D, nx, ny, ndx, ndy, omega, np.float64, args.arith, local_rank, 1,
                                               args.warmup, args.steps, args.rows_per_tile or None)
            lat1.close()
            gbs1 = b.lnx * b.lny * BYTES_PER_CELL / (ms1 / args.steps * 1e-3) / 1e9
            single = {"value": cells * args.steps / (ms1 * 1e-3) / 1e6, "unit": "MLUPS", "ms_per_step": ms1 / args.steps,
                      "achieved_gbs": gbs1, "frac": gbs1 / peak, "kernel": "step_kernel (one pass over HBM per step, 144 B per cell per step)",
                      "same_state_as_temporal_blocking": digest1 == digest_main, "digest": "%016x" % digest1}
        else:
            single = {"value": None, "skipped": "a second %.0f GB lattice does not fit next to nothing: %.0f GB free" % (lattice_b / 1e9, free_b / 1e9)}

    # ---- extras: the other BASELINE.json configurations in the same invocation (so SCALE records them) ---------
    if not args.no_extras and args.workload == "weak16384":
        try:
            # configs[3]: strong scaling, 32768^2 over N x-slabs (one GPU holds it in 155 GB).  The digest is a
            # function of the GLOBAL field only, so equal digests at N = 1, 2, 4, 8 mean bit-identical fields.
            sn = 32768
            latS, msS, digS = timed_lattice(D, sn, sn, args.gpus, 1, omega_for_re(sn), np.float64, args.arith, local_rank, 2, 4, 12)
            bS = latS.blockinfo
            latS.close()
            gbsS = bS.lnx * bS.lny * BYTES_PER_CELL / (msS / 6 * 1e-3) / 1e9
            extra["strong32768"] = {"value": sn * sn * 12 / (msS * 1e-3) / 1e6, "unit": "MLUPS", "ms_per_step": msS / 12, "steps": 12,
                                    "warmup": 4, "blocks": "%dx1 x-slabs" % args.gpus, "achieved_gbs_per_gpu": gbsS, "frac": gbsS / peak,
                                    "digest_after_16_steps": "%016x" % digS,
                                    "note": "the digest depends on the global field only: equal at N=1,2,4,8 <=> bit-identical fields"}
        except Exception as exc:
            extra["strong32768"] = {"value": None, "error": repr(exc)}
        try:
            latF, msF, _ = timed_lattice(D, nx, ny, ndx, ndy, omega, np.float32, args.arith, local_rank, 2, args.warmup, args.steps)
            latF.close()
            gbsF = b.lnx * b.lny * 72 / (msF / (args.steps / 2) * 1e-3) / 1e9
            extra["fp32_temporal_blocking"] = {"value": cells * args.steps / (msF * 1e-3) / 1e6, "unit": "MLUPS", "ms_per_step": msF / args.steps,
                                               "achieved_gbs": gbsF, "frac": gbsF / peak, "bytes_per_cell_per_pass": 72}
        except Exception as exc:
            extra["fp32_temporal_blocking"] = {"value": None, "error": repr(exc)}

    # ---- in-place (AA pattern) variant of the single-step kernel: one buffer instead of the A/B pair (N = 1: one block)
    if not args.no_extras and args.gpus == 1 and single and single.get("value"):
        try:
            latA = lb.Lattice(nx, ny, "cavity", omega=omega, u_wall=0.1, arith=args.arith, devices=local_rank, inplace=True)
            latA.init_equilibrium()
            latA.step(args.warmup)
            latA.sync()
            msA = latA.step_timed(args.steps)
            latA.health()
            digA = latA.checksum()
            latA.close()
            free0 = torch.cuda.mem_get_info(local_rank)[0]
            latB = lb.Lattice(32768, 32768, "cavity", omega=omega_for_re(32768), u_wall=0.1, arith=args.arith, devices=local_rank, inplace=True)
            used = free0 - torch.cuda.mem_get_info(local_rank)[0]
            latB.init_equilibrium()
            latB.step(4)
            latB.sync()
            msB = latB.step_timed(8)
            latB.health()
            latB.close()
            gbsA = cells * BYTES_PER_CELL / (msA / args.steps * 1e-3) / 1e9
            extra["inplace_aa"] = {"value": cells * args.steps / (msA * 1e-3) / 1e6, "unit": "MLUPS", "achieved_gbs": gbsA, "frac": gbsA / peak,
                                   "relative_to_ab_single_step": (cells * args.steps / (msA * 1e-3) / 1e6) / single["value"],
                                   "same_state_as_ab": ("%016x" % digA) == single["digest"],
                                   "lattice_32768_device_gb": used / 1e9, "lattice_32768_mlups": 32768 * 32768 * 8 / (msB * 1e-3) / 1e6,
                                   "note": "ONE copy of the populations (AA pattern, lb_create_ex LB_CREATE_INPLACE), 144 B per cell per step"}
        except Exception as exc:
            extra["inplace_aa"] = {"value": None, "error": repr(exc)}

    # ---- the one workload the reference tree holds timings for: slidingLidMPI.py, 300^2, Re = 1000
    if not args.no_extras and args.gpus == 1:
        try:
            from latticeboltzmann_b200.simulators import sliding_lid_mpi
            sliding_lid_mpi.run(300, 20000, device=local_rank, verbose=False)                     # warm-up
            _, _, secs = sliding_lid_mpi.run(300, 200000, device=local_rank, verbose=False)
            extra["sliding_lid_mpi_300"] = {
                "value": 300 * 300 * 200000 / secs / 1e6, "unit": "MLUPS", "us_per_step": secs / 200000 * 1e6, "steps": 200000,
                "reference_published_mlups": {"1 rank": 4.5, "16 ranks": 95.6, "400 ranks (best)": 237.0},
                "note": "simulators/simple_flows/slidingLidMPI.py (300x300 fluid nodes, Re=1000, uw=0.1; its 10^6-step wall times on bwUniCluster are "
                        "the only timings in the reference tree, amdahldataviewer.py:40-49); here: per-cell boundary table + resident kernel, bit-identical"}
        except Exception as exc:
            extra["sliding_lid_mpi_300"] = {"value": None, "error": repr(exc)}

    cpu = None
    if rank == 0 and args.gpus == 1 and not args.no_cpu_baseline:
        from oracle import opt2_numpy
        v, cores, sample = cpu_reference(nx, ny, omega, 3, 1)
        cap = min(2048, int(os.environ.get("LBM_REF_BLOCK", "2048")))
        one = opt2_numpy.run_independent_blocks(1, cap, cap, omega, 1, 2)
        cpu = {"value": v, "unit": "MLUPS", "cores": cores, "kind": "port", "sample": sample,
               "one_core_value": cap * cap * 2 / one / 1e6,
               "note": "value: the reference's decomposed opt2 run (ghost exchange + np.roll stream + numpy walls + compiled collide) on "
                       "all cores used; one_core_value: one undecomposed %dx%d block on 1 core" % (cap, cap)}
        if strip:
            rows, pre, post = strip
            ok = strip_check_vs_oracle(pre, post, omega, 0.1)
            if parity is not None:
                parity["benchmarked_state_strip"] = {
                    "bit_exact": ok, "rows": list(rows), "columns": "all %d" % b.lny, "steps": 2,
                    "what": "rows of the %dx%d state after the timed region advanced one more pass on the GPU == two oracle steps on the downloaded strip" % (nx, ny)}
                parity["bit_exact"] = parity["bit_exact"] and ok
    if parity is not None and single and single.get("value"):
        parity["two_steps_per_pass_equals_single_steps_at_full_size"] = single["same_state_as_temporal_blocking"]
        parity["bit_exact"] = parity["bit_exact"] and single["same_state_as_temporal_blocking"]

    if rank == 0:
        line = {"metric": METRIC, "value": mlups, "unit": "MLUPS", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": per_step_ms, "higher_is_better": True, "scaling": scaling,
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": desc, "nx": nx, "ny": ny, "ndx": ndx, "ndy": ndy, "omega": omega, "u0": 0.1,
                           "arith": args.arith, "steps_per_hbm_pass": steps_per_pass, "fused_tile_rows": tile_rows,
                           "frame_kernels": ("concurrent with the fused interior kernel (second stream)"
                                             if t2 and os.environ.get("LBM_T2_OVERLAP", "1") != "0" else None),
                           "single_step_roofline_mlups": peak * 1e9 / BYTES_PER_CELL / 1e6,
                           "halo": "in-kernel peer stores over NVLink (CUDA IPC), device-side flags",
                           "numa_bound_cpus": (len(numa) if numa else None),
                           "l2": "inputs larger than L2 (%.1f GB per buffer per GPU, A/B ping-pong)" % (9 * b.lnx * b.lny * 8 / 1e9)},
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches * world,
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": (traffic or {}).get("dram_bytes_per_launch"), "peak_source": peak_src,
                             "bytes_per_cell_per_pass": BYTES_PER_CELL, "steps_per_pass": steps_per_pass,
                             "cells_per_launch": b.lnx * b.lny, "ms_per_pass": per_pass_ms,
                             "frac_per_step_144B": mlups / world * 1e6 * BYTES_PER_CELL / 1e9 / peak,
                             "frac_per_step_144B_note": ("MLUPS per GPU x 144 B / peak, BASELINE's definition: above 1 because temporal blocking moves "
                                                         "144 B per cell per TWO steps; `frac` is per pass (bytes really moved)" if t2 else
                                                         "one pass per step: equals frac"),
                             "kernel": "t2_interior_kernel (+ 2 frame kernels per pass)" if t2 else "step_kernel",
                             "traffic_note": (traffic or {}).get("note")},
                "single_step": single, "parity": parity, "state_digest": ("%016x" % digest_main) if digest_main is not None else None,
                "extra": extra, "cpu_baseline": cpu}
        emit(line)
    import torch.distributed as dist
    if dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
