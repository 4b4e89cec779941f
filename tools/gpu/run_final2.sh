#!/bin/bash
# final one-GPU check of HEAD: whole GPU suite, smoke, both bench arms with default flags
mkdir -p gpurun_out
O=gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 ) > $O/f2_pytest.log
( python __graft_entry__.py smoke 2>&1 | tail -2 ) >> $O/f2_pytest.log
( timeout 600 python bench.py > $O/f2_bench_n1.json 2> $O/f2_bench_n1.err )
( timeout 600 python bench.py --impl reference > $O/f2_bench_ref.json 2> $O/f2_bench_ref.err )
cat $O/f2_pytest.log; tail -3 $O/f2_bench_n1.err | cut -c1-300; cut -c1-400 $O/f2_bench_n1.json; cut -c1-300 $O/f2_bench_ref.json
