#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
( cd tools/ubench && timeout 60 ./chainbench ) > $O/c_chainbench.log 2>&1
for r in 7 0; do
  ( ZC_MSM_TRACE=2 timeout 120 python tools/run_msm.py --rank $r --nranks 8 --prepared --iters 3 2>&1 | tail -30 ) > $O/c_trace_prepared_r$r.log
  ( ZC_MSM_HI_PRIO=0 ZC_MSM_TRACE=2 timeout 120 python tools/run_msm.py --rank $r --nranks 8 --prepared --iters 3 2>&1 | tail -30 ) > $O/c_trace_prepared_loprio_r$r.log
done
( ZC_MSM_TRACE=2 timeout 120 python tools/run_msm.py --rank 0 --nranks 8 --fixed-base --iters 3 2>&1 | tail -20 ) > $O/c_trace_fb_r0.log
( ZC_MSM_TRACE=2 timeout 120 python tools/run_msm.py --rank 7 --nranks 8 --fixed-base --iters 3 2>&1 | tail -20 ) > $O/c_trace_fb_r7.log
( ZC_MSM_TRACE=2 timeout 120 python tools/run_msm.py --rank 3 --nranks 4 --prepared --iters 3 2>&1 | tail -45 ) > $O/c_trace_prepared_r3of4.log
cat $O/c_chainbench.log; tail -25 $O/c_trace_prepared_r7.log; tail -4 $O/c_trace_prepared_loprio_r7.log; tail -14 $O/c_trace_fb_r0.log; tail -2 $O/c_trace_prepared_r0.log $O/c_trace_fb_r7.log $O/c_trace_prepared_r3of4.log
