#!/bin/bash
# batch U: cube2a in 128-thread blocks beside a full accumulation wave, coalesced operand pass, sort above the pass
mkdir -p gpurun_out
O=gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zz_windows.py tests/test_gpu_exchange_local.py -m gpu -x -q -k "msm or exchange" 2>&1 | tail -6 ) > $O/u_pytest.log
: > $O/u_time.log
for mode in --prepared ""; do
  for r in 0 1 2 3 4 5 6 7; do
    ( echo -n "mode[$mode] "; timeout 120 python tools/run_msm.py --rank $r --nranks 8 $mode --iters 6 2>&1 | grep "msm n=" | sort -k7 -n | head -1 ) >> $O/u_time.log
  done
done
for mode in --prepared "" --fixed-base; do ( echo -n "1gpu mode[$mode] "; timeout 120 python tools/run_msm.py $mode --iters 6 2>&1 | grep "msm n=" | sort -k7 -n | head -1 ) >> $O/u_time.log; done
for r in 0 1; do ( ZC_MSM_TRACE=2 timeout 120 python tools/run_msm.py --rank $r --nranks 8 --iters 3 2>&1 | tail -27 ) > $O/u_trace_plain_r$r.log; done
cat $O/u_pytest.log $O/u_time.log; cat $O/u_trace_plain_r0.log
