#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
: > $O/an.log
for lib in "" dusk_zerocaf_b200/libzc_fx0x78.so dusk_zerocaf_b200/libzc_fx0x00.so dusk_zerocaf_b200/libzc_fx0x7b.so dusk_zerocaf_b200/libzc_fx0x70.so; do
  ( echo -n "lib[$lib] "; ZC_LIB_PATH=$lib timeout 200 python tools/time_ops.py fixed 2>&1 | grep "basepoint" | sed 's/.*basepoint/basepoint/' ) >> $O/an.log
done
cat $O/an.log
