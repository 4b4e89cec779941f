#!/bin/bash
# session-3 batch R: wave-filling segment length (14) for one-window accumulations: per-rank timing of the 8-rank decomposition
mkdir -p gpurun_out
O=gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zz_windows.py tests/test_gpu_exchange_local.py -m gpu -x -q -k "msm or exchange" 2>&1 | tail -6 ) > $O/r_pytest.log
for seg in 0 16 14 12 10; do
  for r in 0 3 7; do
    ( echo "== ZC_MSM_SEG=$seg rank $r prepared"; ZC_MSM_SEG=$seg timeout 120 python tools/run_msm.py --rank $r --nranks 8 --prepared --iters 6 2>&1 | grep "msm n=" | sort -k7 -n | head -2 ) >> $O/r_time.log
  done
done
( echo "== default rank 7 plain"; timeout 120 python tools/run_msm.py --rank 7 --nranks 8 --iters 6 2>&1 | grep "msm n=" | sort -k7 -n | head -2
  echo "== seg16 rank 7 plain"; ZC_MSM_SEG=16 timeout 120 python tools/run_msm.py --rank 7 --nranks 8 --iters 6 2>&1 | grep "msm n=" | sort -k7 -n | head -2
  echo "== default 4 ranks rank 3 prepared"; timeout 120 python tools/run_msm.py --rank 3 --nranks 4 --prepared --iters 6 2>&1 | grep "msm n=" | sort -k7 -n | head -2
  echo "== default 1 gpu prepared"; timeout 120 python tools/run_msm.py --prepared --iters 6 2>&1 | grep "msm n=" | sort -k7 -n | head -2 ) >> $O/r_time.log
( ZC_MSM_TRACE=2 timeout 120 python tools/run_msm.py --rank 7 --nranks 8 --prepared --iters 3 2>&1 | tail -26 ) > $O/r_trace_r7.log
cat $O/r_pytest.log $O/r_time.log; tail -25 $O/r_trace_r7.log
