#!/bin/bash
# batch W: which products of the accumulation's addition to inline (alternative library builds, ZC_ACC_INLINE masks)
mkdir -p gpurun_out
O=gpurun_out
: > $O/w_time.log
for lib in "" dusk_zerocaf_b200/libzc_0xff.so dusk_zerocaf_b200/libzc_0xf3.so dusk_zerocaf_b200/libzc_0xf7.so dusk_zerocaf_b200/libzc_0xf4.so; do
  for mode in "--prepared" "--fixed-base --rank 3 --nranks 8" ""; do
    ( echo -n "lib[$lib] mode[$mode] "; ZC_LIB_PATH=$lib timeout 120 python tools/run_msm.py $mode --iters 6 2>&1 | grep "msm n=" | tail -5 | sort -k6 -n | head -1 ) >> $O/w_time.log
  done
done
cat $O/w_time.log
