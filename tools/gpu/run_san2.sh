#!/bin/bash
# compute-sanitizer memcheck over the whole small-size GPU parity suite (final code)
mkdir -p gpurun_out
O=gpurun_out
( timeout 2400 compute-sanitizer --tool memcheck --error-exitcode 77 --launch-timeout 0 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2_surface.py tests/test_gpu_zz_windows.py -m gpu -x -q 2>&1 | tail -12 ) > $O/san2_memcheck.log
echo "exit: $?" >> $O/san2_memcheck.log
( timeout 1200 compute-sanitizer --tool initcheck --error-exitcode 77 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "msm_vs_naive or msm_operand_pass or point_ops or fe_invert or affine or field_ops" 2>&1 | tail -8 ) > $O/san2_initcheck.log
tail -6 $O/san2_memcheck.log; tail -5 $O/san2_initcheck.log
