#!/bin/bash
# session-3 batch O: fast scalar-mul without the unused T products; TMA-staged fixed-base table variants
mkdir -p gpurun_out
O=gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_cpp_surface.py -m gpu -x -q -k "scalar_mul or ristretto or basepoint or point or surface" 2>&1 | tail -6 ) > $O/o_pytest.log
( timeout 200 python tools/time_ops.py smul pt
  ZC_FIXED_LDG=1 timeout 120 python tools/time_ops.py fixed
  for v in 0 1 2; do ZC_FIXED_TMA=$v timeout 120 python tools/time_ops.py fixed; done ) > $O/o_time.log 2>&1
cat $O/o_pytest.log $O/o_time.log
