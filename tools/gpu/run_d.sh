#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
( timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) > $O/d_pytest.log
for r in 7 0; do
  ( ZC_MSM_TRACE=2 timeout 120 python tools/run_msm.py --rank $r --nranks 8 --prepared --iters 3 2>&1 | tail -30 ) > $O/d_trace_prepared_r$r.log
  ( ZC_MSM_FIX_INLINE=9 ZC_MSM_TRACE=2 timeout 120 python tools/run_msm.py --rank $r --nranks 8 --prepared --iters 3 2>&1 | tail -30 ) > $O/d_trace_prepared_fi9_r$r.log
done
( ZC_MSM_TRACE=2 timeout 120 python tools/run_msm.py --rank 0 --nranks 8 --fixed-base --iters 3 2>&1 | tail -20 ) > $O/d_trace_fb_r0.log
( ZC_MSM_TRACE=2 timeout 120 python tools/run_msm.py --rank 7 --nranks 8 --fixed-base --iters 3 2>&1 | tail -20 ) > $O/d_trace_fb_r7.log
( timeout 120 python tools/run_msm.py --prepared --iters 5 --check 2>&1 | tail -8 ) > $O/d_1gpu_prepared.log
( ZC_MSM_STITCH_QUAD=0 timeout 120 python tools/run_msm.py --prepared --iters 5 2>&1 | tail -3 ) > $O/d_1gpu_prepared_oldstitch.log
tail -3 $O/d_pytest.log; tail -24 $O/d_trace_prepared_r7.log; tail -3 $O/d_trace_prepared_fi9_r7.log; tail -14 $O/d_trace_fb_r0.log; tail -n 2 $O/d_trace_prepared_r0.log $O/d_trace_fb_r7.log $O/d_trace_prepared_fi9_r0.log $O/d_1gpu_prepared.log $O/d_1gpu_prepared_oldstitch.log
