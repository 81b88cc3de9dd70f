#!/bin/bash
# round 2: multi-GPU evidence on N GPUs of one box: cross-GPU parity tests (verbose log kept), bench at N, host links
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 1200 python -m pytest tests/test_gpu_multi.py -v > gpurun_out/r02_pytest_multi_n$N.log 2>&1; tail -12 gpurun_out/r02_pytest_multi_n$N.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
tail -c 5500 gpurun_out/r02_bench_n$N.json; tail -3 gpurun_out/r02_bench_n$N.err


