#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
: > $O/al.log
for lib in dusk_zerocaf_b200/libzc_pt0xfc0.so dusk_zerocaf_b200/libzc_pt0xfc3.so dusk_zerocaf_b200/libzc_pt0xff0.so dusk_zerocaf_b200/libzc_pt0xfcf.so dusk_zerocaf_b200/libzc_pt0x3c0.so dusk_zerocaf_b200/libzc_pt0xf0f.so dusk_zerocaf_b200/libzc_pt0xfe0.so; do
  for v in 1 0; do
    ( echo -n "lib[$lib] variant $v: "; ZC_PT_VARIANT=$v ZC_LIB_PATH=$lib timeout 200 python tools/time_ops.py pt 2>&1 | grep "point add" | sed 's/.*point add/point add/' ) >> $O/al.log
  done
done
cat $O/al.log
