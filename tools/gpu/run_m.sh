#!/bin/bash
# session-3 baseline on one GPU: whole GPU test suite, default bench (both arms), MSM timelines
mkdir -p gpurun_out
O=gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > $O/m_pytest.log
( timeout 600 python bench.py > $O/m_bench_n1.json 2> $O/m_bench_n1.err )
( timeout 600 python bench.py --impl reference > $O/m_bench_ref.json 2> $O/m_bench_ref.err )
for r in 7 0; do
  ( ZC_MSM_TRACE=2 timeout 120 python tools/run_msm.py --rank $r --nranks 8 --prepared --iters 3 2>&1 | tail -40 ) > $O/m_trace_prepared_r$r.log
done
( ZC_MSM_TRACE=2 timeout 120 python tools/run_msm.py --rank 0 --nranks 8 --fixed-base --iters 3 2>&1 | tail -30 ) > $O/m_trace_fb_r0.log
( ZC_MSM_TRACE=2 timeout 120 python tools/run_msm.py --prepared --iters 3 2>&1 | tail -120 ) > $O/m_trace_prepared_1gpu.log
cat $O/m_pytest.log; tail -3 $O/m_bench_n1.err | cut -c1-300; cut -c1-1500 $O/m_bench_n1.json; tail -2 $O/m_trace_prepared_r7.log $O/m_trace_fb_r0.log $O/m_trace_prepared_1gpu.log
