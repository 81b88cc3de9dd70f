#!/bin/bash
# 8-GPU box: weak scaling with temporal blocking (default stepping mode)
mkdir -p gpurun_out
run() { n=$1; shift; port=$((29700 + n)); timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n "$@"; }
show() { python -c "
import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[1], d['n_gpus'], d['config']['ndx'], d['config']['ndy'], round(d['value']), round(d['ms_per_step'],4), round(d['roofline']['frac'],4))" $1; }
run 8 --steps 50 --warmup 5 --no-cpu-baseline --no-e2e 2>gpurun_out/weak3_n8.err > gpurun_out/weak3_n8.json; show gpurun_out/weak3_n8.json
run 4 --steps 50 --warmup 5 --no-cpu-baseline --no-e2e 2>gpurun_out/weak3_n4.err > gpurun_out/weak3_n4.json; show gpurun_out/weak3_n4.json
run 8 --workload strong32768 --steps 50 --warmup 5 --no-cpu-baseline --no-e2e 2>gpurun_out/strong3_n8.err > gpurun_out/strong3_n8.json; show gpurun_out/strong3_n8.json
