#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/small_lattices.py 2>&1 | tee gpurun_out/r02_small_lattices_flag.log | cut -c1-150
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; tail -4 gpurun_out/r02_pytest_gpu.log
