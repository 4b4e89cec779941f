#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "basepoint" 2>&1 | tail -4 ) > $O/p_pytest.log
( ZC_FIXED_LDG=1 timeout 120 python tools/time_ops.py fixed
  for v in 0 1 2; do ZC_FIXED_TMA=$v timeout 120 python tools/time_ops.py fixed; done ) > $O/p_time.log 2>&1
cat $O/p_pytest.log $O/p_time.log
