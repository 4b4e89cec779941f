#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
: > $O/aa_time.log
for seg in 0 32 30 28 26 24 20 16; do
  ( echo -n "seg=$seg prepared 1gpu: "; ZC_MSM_SEG=$seg timeout 120 python tools/run_msm.py --prepared --iters 8 2>&1 | grep "msm n=" | tail -6 | sort -k6 -n | head -1 ) >> $O/aa_time.log
done
for seg in 0 28 24; do
  ( echo -n "seg=$seg plain 1gpu: "; ZC_MSM_SEG=$seg timeout 120 python tools/run_msm.py --iters 8 2>&1 | grep "msm n=" | tail -6 | sort -k6 -n | head -1 ) >> $O/aa_time.log
  ( echo -n "seg=$seg prepared rank 1 of 2: "; ZC_MSM_SEG=$seg timeout 120 python tools/run_msm.py --prepared --rank 1 --nranks 2 --iters 8 2>&1 | grep "msm n=" | tail -6 | sort -k6 -n | head -1 ) >> $O/aa_time.log
  ( echo -n "seg=$seg prepared rank 1 of 4: "; ZC_MSM_SEG=$seg timeout 120 python tools/run_msm.py --prepared --rank 1 --nranks 4 --iters 8 2>&1 | grep "msm n=" | tail -6 | sort -k6 -n | head -1 ) >> $O/aa_time.log
done
cat $O/aa_time.log
