#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zz_windows.py tests/test_gpu_fullsize.py -m gpu -x -q -k "msm" 2>&1 | tail -3 ) > $O/ao.log
( timeout 200 python tools/run_msm.py --fixed-base --iters 3 --check 2>&1 | grep -E "fixed-base tables|True|False|msm n=" | tail -4 ) >> $O/ao.log
( timeout 200 python tools/run_msm.py --fixed-base --rank 0 --nranks 8 --iters 3 --check 2>&1 | grep -E "fixed-base tables|True|False" | tail -3 ) >> $O/ao.log
cat $O/ao.log
