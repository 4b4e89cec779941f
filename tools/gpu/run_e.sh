#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > $O/e_pytest.log
( cd tools/ubench && timeout 120 ./pipes5 ) > $O/e_pipes5.log 2>&1
( timeout 120 python __graft_entry__.py smoke 2>&1 | tail -3 ) > $O/e_smoke.log
for r in 7 0; do
  ( ZC_MSM_TRACE=2 timeout 120 python tools/run_msm.py --rank $r --nranks 8 --prepared --iters 3 2>&1 | tail -30 ) > $O/e_trace_prepared_r$r.log
  ( ZC_MSM_SEQ_STAGE1=0 ZC_MSM_TRACE=2 timeout 120 python tools/run_msm.py --rank $r --nranks 8 --prepared --iters 3 2>&1 | tail -30 ) > $O/e_trace_prepared_noseq_r$r.log
done
( timeout 600 python bench.py --steps 10 --warmup 3 > $O/e_bench.json 2> $O/e_bench.err )
( timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $O/e_bench_ref.json 2> $O/e_bench_ref.err )
cat $O/e_pytest.log; cat $O/e_pipes5.log; cat $O/e_smoke.log; tail -3 $O/e_trace_prepared_r7.log $O/e_trace_prepared_noseq_r7.log $O/e_trace_prepared_r0.log $O/e_trace_prepared_noseq_r0.log; tail -5 $O/e_bench.err; head -c 1500 $O/e_bench.json; echo; cat $O/e_bench_ref.json
