#!/bin/bash
# round 2: one-GPU evidence session: GPU suite, default bench, ncu launch list (+ DRAM bytes), ncu --set full of the
# fused kernel, soak and compute-sanitizer over the temporal-blocking paths
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -c 3000 gpurun_out/r02_bench_n1.json; tail -3 gpurun_out/r02_bench_n1.err
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
# launch list of the default bench command (short), with DRAM bytes per launch
timeout 400 ncu --metrics $M --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_weak16384.csv python bench.py --steps 6 --warmup 3 --no-extras --no-e2e --no-cpu-baseline > gpurun_out/r02_ncu_bench.log 2>&1
# full capture of the fused kernel at the bench size
timeout 400 ncu --set full --clock-control none --import-source on -k regex:t2_interior -s 3 -c 1 -o gpurun_out/r02_t2_interior_16384 python tools/profile_target.py 16384 2 10 > gpurun_out/r02_ncu_full.log 2>&1
ncu -i gpurun_out/r02_t2_interior_16384.ncu-rep --page details > gpurun_out/r02_t2_interior_16384_details.txt 2>/dev/null
timeout 300 python tools/soak.py 600 5 2>&1 | tail -2 | tee gpurun_out/r02_soak_overlap.log
for tool in memcheck racecheck; do
  echo "## --tool $tool" >> gpurun_out/r02_sanitizer_overlap.txt
  timeout 300 compute-sanitizer --tool $tool python tools/sanitize_target.py temporal_tma temporal_tma_f32 temporal_blocks_2x2 2>&1 | grep -E "^ok|all paths|SUMMARY|COMPUTE-SANITIZER|Error|error" >> gpurun_out/r02_sanitizer_overlap.txt
done
cat gpurun_out/r02_sanitizer_overlap.txt
ls -la gpurun_out | head -40
