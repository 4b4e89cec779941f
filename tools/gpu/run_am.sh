#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_cpp_surface.py -m gpu -x -q -k "point or cfg3 or surface or kats or operator" 2>&1 | tail -3 ) > $O/am.log
( timeout 200 python tools/time_ops.py pt ) >> $O/am.log 2>&1
cat $O/am.log
