#!/bin/bash
# session-3 batch Q: interleaved Montgomery squaring -- micro-benchmark, the whole GPU test suite, scalar-mul timing
mkdir -p gpurun_out
O=gpurun_out
( cd tools/ubench && timeout 120 ./sqrbench ) > $O/q_sqrbench.log 2>&1
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > $O/q_pytest.log
( timeout 200 python tools/time_ops.py smul ) > $O/q_time.log 2>&1
cat $O/q_sqrbench.log $O/q_pytest.log $O/q_time.log
