#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3
python bench.py --steps 50 --warmup 5 > gpurun_out/bench_weak16384_exact.json 2>gpurun_out/bench_b.err; tail -c 400 gpurun_out/bench_weak16384_exact.json; echo
python bench.py --steps 50 --warmup 5 --arith fast --no-cpu-baseline --no-e2e > gpurun_out/bench_weak16384_fast.json 2>>gpurun_out/bench_b.err
python bench.py --workload cavity4096 --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/bench_cavity4096_exact.json 2>>gpurun_out/bench_b.err
python bench.py --workload cavity4096 --steps 200 --warmup 10 --arith fast --no-cpu-baseline --no-e2e > gpurun_out/bench_cavity4096_fast.json 2>>gpurun_out/bench_b.err
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
for a in exact fast; do
ncu --metrics $M --clock-control none -k regex:step_kernel -s 5 -c 3 --csv --log-file gpurun_out/ncu_dram_16384_$a.csv python bench.py --steps 5 --warmup 3 --arith $a --no-e2e --no-cpu-baseline > /dev/null 2>&1
ncu --metrics $M --clock-control none -k regex:step_kernel -s 5 -c 3 --csv --log-file gpurun_out/ncu_dram_4096_$a.csv python bench.py --workload cavity4096 --steps 5 --warmup 3 --arith $a --no-e2e --no-cpu-baseline > /dev/null 2>&1
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_weak16384.csv python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 6 -c 1 -o gpurun_out/prof_step_exact_v2 python bench.py --workload cavity4096 --steps 6 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
python - <<'PY'
import sys; sys.path.insert(0,'.')
from tools.quick_bench import run
for dt in ("float32",):
    for ar in ("exact","fast"):
        for rows in (4,8,16):
            print(dt, ar, rows, run(4096, dt, ar, rows), flush=True)
PY
ls gpurun_out | head -40
