#!/bin/bash
run2() { port=$((29600 + RANDOM % 300)); timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | grep '^{' | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['config']['ndx'], d['config']['ndy'], round(d['value']), round(d['ms_per_step'],4))"; }
echo "2x1 normal"; run2
echo "1x2 normal"; LBM_BENCH_GRID=1x2 run2
echo "1x2 skip y pushes"; LBM_BENCH_GRID=1x2 LBM_NATIVE_LIB=$PWD/tools/dbg/liblbm_skipy.so run2
