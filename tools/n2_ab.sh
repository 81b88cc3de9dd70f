mkdir -p gpurun_out
for rep in 1 2; do
for cfg in "1 0" "0 0" "1 32" "0 32"; do set -- $cfg
  echo "overlap=$1 rows=$2 (0: default)" | tee -a gpurun_out/v8_n2_ab.log
  LBM_T2_OVERLAP=$1 LBM_T2_ROWS=$2 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 50 --warmup 5 --no-extras --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['clocks'])" | tee -a gpurun_out/v8_n2_ab.log
done; done
