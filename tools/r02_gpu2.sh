#!/bin/bash
# round 2, call 2: TMA-staged variants of the fused kernel vs the compact-ring baseline
mkdir -p gpurun_out
timeout 900 python tools/t2_variants.py run 16384 20 tma3 tma4 2>&1 | tee gpurun_out/r02_variants3.log
