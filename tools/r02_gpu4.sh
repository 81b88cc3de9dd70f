#!/bin/bash
# round 2: resident multi-step kernel -- parity on both stepping paths, then timing of configs[0] and configs[1]
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; tail -5 gpurun_out/r02_pytest_gpu.log
for v in res256 res512 res1024; do
  echo "== $v"; LBM_NATIVE_LIB=latticeboltzmann_b200/csrc/variants/lib_$v.so timeout 600 python tools/small_lattices.py 2>&1 | tee -a gpurun_out/r02_small_lattices_$v.log
done
