#!/bin/bash
# round-2 ncu captures: one --set full capture per kernel (second launch of each, warm), plus the launch list of the bench
mkdir -p gpurun_out
O=gpurun_out
cap() {  # name regex skip
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c 1 -f -o $O/r02_$1 python tools/run_kernels.py > $O/r02_$1.log 2>&1
}
cap fe_mul_square '^fe_mul_square_kernel' 1
cap fe_mul_square_packed 'fe_mul_square_packed_kernel' 1
cap pt_add 'pt_op_kernel' 1
cap scalar_mul_strict 'scalar_mul_strict_kernel' 1
cap scalar_mul_fast 'scalar_mul_fast_kernel' 1
cap msm_accum 'msm_accum_kernel' 4
cap msm_chain 'msm_chain_kernel' 4
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --skip-cpu --no-sustain > $O/r02_bench_under_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_kernels.csv python tools/run_kernels.py > /dev/null 2>&1
ls -la $O/r02_*
