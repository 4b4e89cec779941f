#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
for r in 0 1; do
  ( ZC_MSM_TRACE=2 timeout 120 python tools/run_msm.py --rank $r --nranks 8 --iters 3 2>&1 | tail -34 ) > $O/t_trace_plain_r$r.log
  ( ZC_MSM_TRACE=2 timeout 120 python tools/run_msm.py --rank $r --nranks 8 --fixed-base --iters 3 2>&1 | tail -16 ) > $O/t_trace_fb_r$r.log
done
( ZC_MSM_TRACE=2 timeout 120 python tools/run_msm.py --rank 0 --nranks 8 --prepared --iters 3 2>&1 | tail -30 ) > $O/t_trace_prep_r0.log
tail -n 40 $O/t_trace_plain_r0.log $O/t_trace_plain_r1.log $O/t_trace_fb_r0.log $O/t_trace_fb_r1.log $O/t_trace_prep_r0.log
