#!/bin/bash
# session-3 batch N: TMA-staged fixed-base table vs the L1 path, point-op block shapes
mkdir -p gpurun_out
O=gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "basepoint or point or ristretto" 2>&1 | tail -6 ) > $O/n_pytest.log
( for v in 0 1 2 3; do ZC_PT_VARIANT=$v timeout 120 python tools/time_ops.py pt; done
  ZC_FIXED_LDG=1 timeout 120 python tools/time_ops.py fixed
  timeout 120 python tools/time_ops.py fixed ) > $O/n_time.log 2>&1
cat $O/n_pytest.log $O/n_time.log
