#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
: > $O/ak.log
for lib in "" dusk_zerocaf_b200/libzc_smf5.so dusk_zerocaf_b200/libzc_smf4.so; do
  ( echo -n "lib[$lib] "; ZC_LIB_PATH=$lib timeout 200 python tools/time_ops.py smul 2>&1 | grep "mode 1" ) >> $O/ak.log
done
cat $O/ak.log
