#!/bin/bash
# per-rank times of the 8-rank decomposition on one GPU (prepared, plain, fixed-base), current defaults
mkdir -p gpurun_out
O=gpurun_out
: > $O/s_time.log
for mode in --prepared "" --fixed-base; do
  for r in 0 1 2 3 4 5 6 7; do
    ( echo -n "mode[$mode] "; timeout 120 python tools/run_msm.py --rank $r --nranks 8 $mode --iters 6 2>&1 | grep "msm n=" | sort -k7 -n | head -1 ) >> $O/s_time.log
  done
done
cat $O/s_time.log
