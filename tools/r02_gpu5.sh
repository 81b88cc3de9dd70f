#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; tail -15 gpurun_out/r02_pytest_gpu.log
timeout 600 python tools/aa_bench.py 2>&1 | tee gpurun_out/r02_aa_bench.log | tail -3
