#!/bin/bash
python tools/quick_bench4.py
for v in v_pref v_rows2 v_rows2_minb3; do LBM_NATIVE_LIB=$PWD/tools/dbg/lib_$v.so python tools/quick_bench4.py; done
