#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
( ZC_SMF_SMEM=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "scalar_mul or ristretto_vectors" 2>&1 | tail -3 ) > $O/ac_pytest.log
( timeout 200 python tools/time_ops.py smul | grep "mode 1"; ZC_SMF_SMEM=1 timeout 300 python tools/time_ops.py smul | grep "mode 1" ) > $O/ac_time.log 2>&1
cat $O/ac_pytest.log $O/ac_time.log
