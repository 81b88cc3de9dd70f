#!/bin/bash
# first measured pass: smoke, bench lines, ncu launch list + full capture of the step kernel
set -x
mkdir -p gpurun_out
python __graft_entry__.py --smoke 2>&1 | tail -3
python bench.py --steps 50 --warmup 5 > gpurun_out/bench_weak16384_n1.json 2> gpurun_out/bench_weak16384_n1.err; tail -c 3000 gpurun_out/bench_weak16384_n1.json; tail -5 gpurun_out/bench_weak16384_n1.err
python bench.py --workload cavity4096 --steps 200 --warmup 10 > gpurun_out/bench_cavity4096.json 2>gpurun_out/bench_cavity4096.err; tail -c 2500 gpurun_out/bench_cavity4096.json
python bench.py --workload cavity4096 --arith exact --steps 200 --warmup 10 --no-e2e --no-cpu-baseline > gpurun_out/bench_cavity4096_exact.json 2>&1; tail -c 1500 gpurun_out/bench_cavity4096_exact.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_cavity4096.csv python bench.py --workload cavity4096 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
tail -5 gpurun_out/launches_cavity4096.csv
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 6 -c 2 -o gpurun_out/prof_step_fast python bench.py --workload cavity4096 --steps 6 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_fast.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 6 -c 2 -o gpurun_out/prof_step_exact python bench.py --workload cavity4096 --arith exact --steps 6 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_exact.log 2>&1
ls -la gpurun_out
