#!/bin/bash
# batch V: persistent operand pass beside the high-priority sort
mkdir -p gpurun_out
O=gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zz_windows.py -m gpu -x -q -k "msm" 2>&1 | tail -6 ) > $O/v_pytest.log
: > $O/v_time.log
for r in 0 1 4 5 7; do
  ( echo -n "plain "; timeout 120 python tools/run_msm.py --rank $r --nranks 8 --iters 6 2>&1 | grep "msm n=" | tail -5 | sort -k6 -n | head -1 ) >> $O/v_time.log
done
for mode in --prepared "" --fixed-base; do ( echo -n "1gpu mode[$mode] "; timeout 120 python tools/run_msm.py $mode --iters 6 2>&1 | grep "msm n=" | tail -5 | sort -k6 -n | head -1 ) >> $O/v_time.log; done
( ZC_MSM_TRACE=2 timeout 120 python tools/run_msm.py --rank 0 --nranks 8 --iters 3 2>&1 | tail -27 ) > $O/v_trace_plain_r0.log
cat $O/v_pytest.log $O/v_time.log; head -12 $O/v_trace_plain_r0.log
