#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
: > $O/ab_time.log
for seg in 0 14 20 24 28 32; do
  ( echo -n "seg=$seg fb rank 3 of 8: "; ZC_MSM_SEG=$seg timeout 120 python tools/run_msm.py --fixed-base --rank 3 --nranks 8 --iters 8 2>&1 | grep "msm n=" | tail -6 | awk '{print $6}' | sort -n | head -1 ) >> $O/ab_time.log
done
for seg in 0 28; do
  ( echo -n "seg=$seg fb rank 0 of 8: "; ZC_MSM_SEG=$seg timeout 120 python tools/run_msm.py --fixed-base --rank 0 --nranks 8 --iters 8 2>&1 | grep "msm n=" | tail -6 | awk '{print $6}' | sort -n | head -1 ) >> $O/ab_time.log
done
cat $O/ab_time.log
