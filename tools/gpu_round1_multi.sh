#!/bin/bash
# 8-GPU box: weak/strong scaling series (multi-GPU parity: tests/test_gpu_multi.py)
mkdir -p gpurun_out
run() { n=$1; shift; port=$((29600 + RANDOM % 300)); if [ "$n" = 1 ]; then timeout 600 python bench.py --gpus 1 "$@"; else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n "$@"; fi; }
show() { python -c "
import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[1], d['n_gpus'], round(d['value']), round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d['e2e'] and d['e2e'].get('value'))" $1; }
for n in 8 4 2 1; do
  run $n --steps 50 --warmup 5 --no-cpu-baseline 2>gpurun_out/weak_n$n.err | grep '^{' > gpurun_out/weak_n$n.json; show gpurun_out/weak_n$n.json
done
for n in 8 4 2; do
  run $n --workload strong32768 --steps 50 --warmup 5 --no-cpu-baseline --no-e2e 2>gpurun_out/strong_n$n.err | grep '^{' > gpurun_out/strong_n$n.json; show gpurun_out/strong_n$n.json
done
run 1 --workload strong32768 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>gpurun_out/strong_n1.err | grep '^{' > gpurun_out/strong_n1.json; show gpurun_out/strong_n1.json
for f in gpurun_out/*.err; do echo "== $f"; grep -v "OMP_NUM_THREADS\|^\*\*\*\*\|^$" $f | tail -n 3; done
