#!/bin/bash
# final one-GPU batch: whole GPU test suite, both bench arms with default flags, ncu captures + launch lists
mkdir -p gpurun_out
O=gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > $O/f_pytest.log
( timeout 600 python bench.py > $O/f_bench_n1.json 2> $O/f_bench_n1.err )
( timeout 600 python bench.py --impl reference > $O/f_bench_ref.json 2> $O/f_bench_ref.err )
bash tools/gpu/run_ncu.sh > $O/f_ncu.log 2>&1
( ZC_MSM_TRACE=2 timeout 120 python tools/run_msm.py --prepared --iters 3 2>&1 | tail -60 ) > $O/f_trace_prepared_1gpu.log
cat $O/f_pytest.log; tail -3 $O/f_bench_n1.err | cut -c1-300; cut -c1-600 $O/f_bench_n1.json; tail -4 $O/f_ncu.log
