#!/bin/bash
python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -3
python tools/quick_bench3.py 4096
for v in v_minb4 v_minb2 v_pref3 v_pref2; do LBM_NATIVE_LIB=$PWD/tools/dbg/lib_$v.so python tools/quick_bench3.py 4096; done
