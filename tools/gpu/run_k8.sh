#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zz_windows.py tests/test_gpu_fullsize.py -m gpu -x -q -k "msm" 2>&1 | tail -8 ) > $O/k8_pytest.log
for sp in 1 2 3 4; do
  export ZC_MSM_SPLIT=$sp
  ( timeout 120 python tools/run_msm.py --rank 0 --nranks 8 --fixed-base --iters 6 --check 2>&1 | tail -5 ) > $O/k8_fb_r0_split$sp.log
  ( timeout 120 python tools/run_msm.py --rank 5 --nranks 8 --fixed-base --iters 6 2>&1 | tail -3 ) > $O/k8_fb_r5_split$sp.log
done
ZC_MSM_SPLIT=2 ZC_MSM_TRACE=2 timeout 120 python tools/run_msm.py --rank 0 --nranks 8 --fixed-base --iters 3 2>&1 | tail -24 > $O/k8_trace_fb_r0_split2.log
ZC_MSM_SPLIT=4 ZC_MSM_TRACE=2 timeout 120 python tools/run_msm.py --rank 0 --nranks 8 --fixed-base --iters 3 2>&1 | tail -40 > $O/k8_trace_fb_r0_split4.log
unset ZC_MSM_SPLIT
( timeout 120 python tools/run_msm.py --rank 0 --nranks 2 --fixed-base --iters 4 --check 2>&1 | tail -4 ) > $O/k8_fb_r0of2.log
( timeout 120 python tools/run_msm.py --fixed-base --iters 4 --check 2>&1 | tail -4 ) > $O/k8_fb_1gpu.log
cat $O/k8_pytest.log
for f in $O/k8_fb*.log; do echo "$f: $(grep 'msm n=' $f | sort -k7 -n | head -1) $(grep -c True $f)"; done
cat $O/k8_trace_fb_r0_split2.log
