#!/bin/bash
# final one-GPU check of HEAD: whole GPU suite, smoke, both bench arms with default flags, ncu of the point-add kernel
mkdir -p gpurun_out
O=gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 ) > $O/f3_pytest.log
( python __graft_entry__.py smoke 2>&1 | tail -2 ) >> $O/f3_pytest.log
( timeout 600 python bench.py > $O/f3_bench_n1.json 2> $O/f3_bench_n1.err )
( timeout 600 python bench.py --impl reference > $O/f3_bench_ref.json 2> $O/f3_bench_ref.err )
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:pt_op_kernel" -s 1 -c 1 -f -o $O/r02_pt_add python tools/run_kernels.py > $O/r02_pt_add.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --skip-cpu --no-sustain > $O/r02_bench_under_ncu.log 2>&1
cat $O/f3_pytest.log; tail -3 $O/f3_bench_n1.err | cut -c1-300; cut -c1-300 $O/f3_bench_n1.json
