#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
for k in place fine count; do
  ZC_MSM_TRACE=1 timeout 300 ncu --set full --clock-control none --import-source on -k "regex:msm_sort_${k}_kernel" -s 4 -c 1 -f -o $O/k7_$k python tools/run_msm.py --rank 0 --nranks 8 --fixed-base --iters 2 > $O/k7_$k.log 2>&1
done
ls -la $O/k7_*
