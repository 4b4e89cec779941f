#!/bin/bash
# round-2 GPU batch B: after the branch-free / lazy / small-constant quad operations
mkdir -p gpurun_out
O=gpurun_out
( timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) > $O/b_pytest.log
( cd tools/ubench && timeout 60 ./chainbench ) > $O/b_chainbench.log 2>&1
for r in 7 0; do
  ( ZC_MSM_TRACE=2 timeout 120 python tools/run_msm.py --rank $r --nranks 8 --prepared --iters 3 --check 2>&1 | tail -30 ) > $O/b_trace_prepared_r$r.log
done
( ZC_MSM_TRACE=2 timeout 120 python tools/run_msm.py --rank 0 --nranks 8 --fixed-base --iters 3 --check 2>&1 | tail -20 ) > $O/b_trace_fb_r0.log
( timeout 120 python tools/run_msm.py --prepared --iters 5 --check 2>&1 | tail -8 ) > $O/b_1gpu_prepared.log
tail -3 $O/b_pytest.log; cat $O/b_chainbench.log; tail -25 $O/b_trace_prepared_r7.log; tail -16 $O/b_trace_fb_r0.log; cat $O/b_1gpu_prepared.log
