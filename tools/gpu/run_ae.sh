#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
: > $O/ae_time.log
( ZC_MSM_ACC_OVERLAP=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zz_windows.py -m gpu -x -q -k "msm" 2>&1 | tail -3 ) >> $O/ae_time.log
for ov in 0 1; do
  for mode in "--prepared" "" "--prepared --rank 1 --nranks 2" "--prepared --rank 1 --nranks 4" "--prepared --rank 5 --nranks 8" "--prepared --rank 0 --nranks 8"; do
    ( echo -n "overlap=$ov [$mode]: "; ZC_MSM_ACC_OVERLAP=$ov timeout 120 python tools/run_msm.py $mode --iters 8 2>&1 | grep "msm n=" | tail -6 | awk '{print $6}' | sort -n | head -1 ) >> $O/ae_time.log
  done
done
( ZC_MSM_ACC_OVERLAP=1 ZC_MSM_TRACE=2 timeout 120 python tools/run_msm.py --prepared --iters 3 2>&1 | tail -52 | grep -E "accum|chain|sort_fine|msm n=" ) > $O/ae_trace.log
cat $O/ae_time.log $O/ae_trace.log
