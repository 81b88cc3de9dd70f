#!/bin/bash
# round 2, call 1: GPU suite with the cp.async-staged fused kernel, variant sweep, ncu of the round-1 and the new kernel
mkdir -p gpurun_out
V=latticeboltzmann_b200/csrc/variants
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu1.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu1.log
timeout 1200 python tools/t2_variants.py run 16384 20 2>&1 | tee gpurun_out/r02_variants1.log
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
for v in r1 async4; do
  LBM_NATIVE_LIB=$V/lib_$v.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:t2_interior -s 2 -c 1 -o gpurun_out/r02_t2_$v python tools/profile_target.py 8192 2 8 > gpurun_out/r02_ncu_$v.log 2>&1
  ncu -i gpurun_out/r02_t2_$v.ncu-rep --page raw --csv > gpurun_out/r02_t2_${v}_raw.csv 2>/dev/null
  LBM_NATIVE_LIB=$V/lib_$v.so timeout 600 ncu --metrics $M --clock-control none -k regex:t2_ -s 6 -c 6 --csv --log-file gpurun_out/r02_dram_16384_$v.csv python tools/profile_target.py 16384 2 8 > /dev/null 2>&1
done
ls -la gpurun_out | head -30
