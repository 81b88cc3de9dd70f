#!/bin/bash
python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -2
p=29611
for g in 2x1 1x2; do
  p=$((p+1))
  LBM_BENCH_GRID=$g python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $p bench.py --gpus 2 --steps 50 --warmup 5 --no-e2e 2>gpurun_out/t2_2gpu_$g.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['config']['ndx'], d['config']['ndy'], round(d['value']), d['ms_per_step'], d['roofline']['frac'])" || tail -5 gpurun_out/t2_2gpu_$g.err
done
