#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
: > $O/ad_time.log
for sp in 1 0; do
  for mode in "--prepared" "" "--prepared --rank 1 --nranks 2"; do
    ( echo -n "split=$sp [$mode]: "; ZC_MSM_GROUP_SPLIT=$sp timeout 120 python tools/run_msm.py $mode --iters 8 2>&1 | grep "msm n=" | tail -6 | awk '{print $6}' | sort -n | head -1 ) >> $O/ad_time.log
  done
done
( timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zz_windows.py -m gpu -x -q -k "msm" 2>&1 | tail -3 ) >> $O/ad_time.log
( ZC_MSM_TRACE=2 timeout 120 python tools/run_msm.py --prepared --iters 3 2>&1 | tail -52 ) > $O/ad_trace.log
cat $O/ad_time.log
