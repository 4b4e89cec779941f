#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zz_windows.py tests/test_gpu_exchange_local.py -m gpu -x -q -k "msm or exchange" 2>&1 | tail -4 ) > $O/z_pytest.log
: > $O/z_time.log
for mode in --prepared ""; do
  for r in 0 1 2 3 4 5 6 7; do
    ( echo -n "mode[$mode] "; timeout 120 python tools/run_msm.py --rank $r --nranks 8 $mode --iters 6 2>&1 | grep "msm n=" | tail -5 | sort -k6 -n | head -1 ) >> $O/z_time.log
  done
done
for mode in --prepared "" --fixed-base; do ( echo -n "1gpu mode[$mode] "; timeout 120 python tools/run_msm.py $mode --iters 6 2>&1 | grep "msm n=" | tail -5 | sort -k6 -n | head -1 ) >> $O/z_time.log; done
( ZC_MSM_TRACE=2 timeout 120 python tools/run_msm.py --rank 0 --nranks 8 --prepared --iters 3 2>&1 | tail -26 ) > $O/z_trace_prep_r0.log
cat $O/z_pytest.log $O/z_time.log $O/z_trace_prep_r0.log
