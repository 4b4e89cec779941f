#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
( ZC_STRICT_DBL=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -k "scalar_mul_strict or ristretto_vectors or cfg4" 2>&1 | tail -3 ) > $O/aj.log
( timeout 200 python tools/time_ops.py smul | grep "mode 0"; ZC_STRICT_DBL=1 timeout 200 python tools/time_ops.py smul | grep "mode 0" ) >> $O/aj.log 2>&1
cat $O/aj.log
