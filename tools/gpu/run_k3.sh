#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
M=gpu__time_duration.sum,launch__waves_per_multiprocessor,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum
ZC_MSM_TRACE=1 timeout 300 ncu --metrics $M --clock-control none -c 60 --csv --log-file $O/k3_fb_r0.csv python tools/run_msm.py --rank 0 --nranks 8 --fixed-base --iters 2 > $O/k3_fb_r0.log 2>&1
ZC_MSM_TRACE=1 timeout 300 ncu --metrics $M --clock-control none -c 90 --csv --log-file $O/k3_prep_r7.csv python tools/run_msm.py --rank 7 --nranks 8 --prepared --iters 2 > $O/k3_prep_r7.log 2>&1
tail -3 $O/k3_fb_r0.log
