#!/bin/bash
# batch X: fast scalar-mul with inlined output products (ZC_SMF_INLINE 1 / 3) vs out-of-line; full GPU test suite on the default build
mkdir -p gpurun_out
O=gpurun_out
: > $O/x_time.log
for lib in "" dusk_zerocaf_b200/libzc_smf1.so dusk_zerocaf_b200/libzc_smf3.so; do
  ( echo "lib[$lib]"; ZC_LIB_PATH=$lib timeout 200 python tools/time_ops.py smul 2>&1 | grep "mode 1" ) >> $O/x_time.log
done
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 ) > $O/x_pytest.log
cat $O/x_time.log $O/x_pytest.log
