#!/bin/bash
# round 2: compute-sanitizer over every stepping path (small lattices)
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_target.py > gpurun_out/r02_sanitizer_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "^ok|all paths|ERROR SUMMARY|RACECHECK SUMMARY|hazard|Error:|error" gpurun_out/r02_sanitizer_$tool.log | sort | uniq -c | head -30
done
