#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zz_windows.py tests/test_gpu_fullsize.py -m gpu -x -q -k "msm" 2>&1 | tail -8 ) > $O/k6_pytest.log
( ZC_MSM_TRACE=2 timeout 120 python tools/run_msm.py --rank 0 --nranks 8 --fixed-base --iters 3 2>&1 | tail -16 ) > $O/k6_trace_fb_r0.log
( ZC_MSM_TRACE=2 timeout 120 python tools/run_msm.py --rank 7 --nranks 8 --prepared --iters 3 2>&1 | tail -28 ) > $O/k6_trace_prep_r7.log
( timeout 120 python tools/run_msm.py --rank 0 --nranks 8 --fixed-base --iters 5 2>&1 | tail -3 ) > $O/k6_fb_r0.log
( timeout 120 python tools/run_msm.py --rank 7 --nranks 8 --prepared --iters 5 2>&1 | tail -3 ) > $O/k6_prep_r7.log
( timeout 120 python tools/run_msm.py --rank 0 --nranks 8 --prepared --iters 5 2>&1 | tail -3 ) > $O/k6_prep_r0.log
( timeout 120 python tools/run_msm.py --prepared --iters 5 --check 2>&1 | tail -4 ) > $O/k6_1gpu_prep.log
( ZC_MSM_TRACE=2 timeout 120 python tools/run_msm.py --prepared --iters 2 2>&1 | tail -75 ) > $O/k6_trace_1gpu_prep.log
( timeout 120 python tools/run_msm.py --iters 5 2>&1 | tail -3 ) > $O/k6_1gpu_plain.log
( timeout 120 python tools/run_msm.py --c 12 --iters 3 --check 2>&1 | tail -3 ) > $O/k6_1gpu_c12.log
cat $O/k6_pytest.log
for f in $O/k6_*.log; do echo "$f: $(tail -n 1 $f)"; done
tail -14 $O/k6_trace_fb_r0.log
