#!/bin/bash
# new two-level counting sort: MSM tests, A/B against the atomic sort, timelines
mkdir -p gpurun_out
O=gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zz_windows.py tests/test_gpu_fullsize.py -m gpu -x -q -k "msm" 2>&1 | tail -8 ) > $O/k_pytest.log
for mode in counting atomic; do
  export ZC_MSM_SORT=$mode
  ( ZC_MSM_TRACE=2 timeout 120 python tools/run_msm.py --rank 0 --nranks 8 --fixed-base --iters 3 2>&1 | tail -16 ) > $O/k_trace_fb_r0_$mode.log
  ( ZC_MSM_TRACE=2 timeout 120 python tools/run_msm.py --rank 7 --nranks 8 --prepared --iters 3 2>&1 | tail -28 ) > $O/k_trace_prep_r7_$mode.log
  ( timeout 120 python tools/run_msm.py --rank 0 --nranks 8 --fixed-base --iters 5 2>&1 | tail -3 ) > $O/k_fb_r0_$mode.log
  ( timeout 120 python tools/run_msm.py --rank 7 --nranks 8 --prepared --iters 5 2>&1 | tail -3 ) > $O/k_prep_r7_$mode.log
  ( timeout 120 python tools/run_msm.py --rank 0 --nranks 8 --prepared --iters 5 2>&1 | tail -3 ) > $O/k_prep_r0_$mode.log
  ( timeout 120 python tools/run_msm.py --prepared --iters 5 --check 2>&1 | tail -4 ) > $O/k_1gpu_prep_$mode.log
  ( timeout 120 python tools/run_msm.py --iters 5 2>&1 | tail -3 ) > $O/k_1gpu_plain_$mode.log
  ( timeout 120 python tools/run_msm.py --c 12 --iters 3 --check 2>&1 | tail -3 ) > $O/k_1gpu_c12_$mode.log
done
cat $O/k_pytest.log
for f in $O/k_*_counting.log $O/k_*_atomic.log; do echo "$f: $(tail -n 1 $f)"; done
cat $O/k_trace_fb_r0_counting.log
