#!/bin/bash
# compute-sanitizer (memcheck, then racecheck on shared memory) over the small-size MSM / scalar-mul / fixed-base tests
mkdir -p gpurun_out
O=gpurun_out
( timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "msm_operand_pass" 2>&1 | tail -3 ) > $O/san_newtest.log
( timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 77 --launch-timeout 0 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "msm_vs_naive or msm_edges or msm_sharding or msm_operand_pass or msm_fixed_base or msm_prepared or scalar_mul_fast or fe_invert or basepoint" 2>&1 | tail -25 ) > $O/san_memcheck.log
echo "memcheck exit: $?" >> $O/san_memcheck.log
( ZC_FIXED_TMA=1 timeout 600 compute-sanitizer --tool memcheck --error-exitcode 77 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "basepoint" 2>&1 | tail -8 ) > $O/san_memcheck_tma.log
( timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 77 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "msm_vs_naive or msm_sharding or msm_operand_pass" 2>&1 | tail -25 ) > $O/san_racecheck.log
cat $O/san_newtest.log; tail -12 $O/san_memcheck.log; tail -6 $O/san_memcheck_tma.log; tail -14 $O/san_racecheck.log
