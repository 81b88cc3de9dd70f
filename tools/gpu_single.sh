#!/bin/bash
# round 2: full GPU suite, default bench, launch list + dram bytes + ncu --set full of the shipped kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; tail -5 gpurun_out/r02_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -c 6000 gpurun_out/r02_bench_n1.json; tail -5 gpurun_out/r02_bench_n1.err
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
# launch list of the default bench command (short), with DRAM bytes per launch
timeout 900 ncu --metrics $M --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_weak16384.csv python bench.py --steps 6 --warmup 3 --no-extras --no-e2e --no-cpu-baseline > gpurun_out/r02_ncu_bench.log 2>&1
timeout 900 ncu --metrics $M --clock-control none -c 40 --csv --log-file gpurun_out/r02_launches_weak16384_t1.csv python bench.py --steps 6 --warmup 3 --temporal 1 --no-extras --no-e2e --no-cpu-baseline >> gpurun_out/r02_ncu_bench.log 2>&1
# full capture of the fused kernel at the bench size
timeout 900 ncu --set full --clock-control none --import-source on -k regex:t2_interior -s 3 -c 1 -o gpurun_out/r02_t2_interior_16384 python tools/profile_target.py 16384 2 10 > gpurun_out/r02_ncu_full.log 2>&1
ncu -i gpurun_out/r02_t2_interior_16384.ncu-rep --page raw --csv > gpurun_out/r02_t2_interior_16384_raw.csv 2>/dev/null
ncu -i gpurun_out/r02_t2_interior_16384.ncu-rep --page details > gpurun_out/r02_t2_interior_16384_details.txt 2>/dev/null
ls -la gpurun_out | head -40
