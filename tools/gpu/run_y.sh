#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zz_windows.py -m gpu -x -q -k "msm" 2>&1 | tail -4 ) > $O/y_pytest.log
: > $O/y_time.log
for r in 0 1 7; do ( echo -n "fb "; timeout 120 python tools/run_msm.py --rank $r --nranks 8 --fixed-base --iters 6 2>&1 | grep "msm n=" | tail -5 | sort -k6 -n | head -1 ) >> $O/y_time.log; done
for r in 0 5 7; do ( echo -n "prepared "; timeout 120 python tools/run_msm.py --rank $r --nranks 8 --prepared --iters 6 2>&1 | grep "msm n=" | tail -5 | sort -k6 -n | head -1 ) >> $O/y_time.log; done
( ZC_MSM_TRACE=2 timeout 120 python tools/run_msm.py --rank 0 --nranks 8 --fixed-base --iters 3 2>&1 | tail -14 ) > $O/y_trace_fb_r0.log
( ZC_MSM_TRACE=2 timeout 120 python tools/run_msm.py --rank 5 --nranks 8 --prepared --iters 3 2>&1 | tail -26 ) > $O/y_trace_prep_r5.log
cat $O/y_pytest.log $O/y_time.log $O/y_trace_fb_r0.log $O/y_trace_prep_r5.log
