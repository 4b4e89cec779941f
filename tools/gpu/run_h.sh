#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zz_windows.py tests/test_gpu_fullsize.py -m gpu -x -q -k "msm" 2>&1 | tail -5 ) > $O/h_pytest.log
for r in 7 3 0; do
  ( ZC_MSM_TRACE=2 timeout 120 python tools/run_msm.py --rank $r --nranks 8 --prepared --iters 3 2>&1 | tail -24 ) > $O/h_trace_prepared_r$r.log
  ( ZC_MSM_ACC_TPB=128 ZC_MSM_TRACE=2 timeout 120 python tools/run_msm.py --rank $r --nranks 8 --prepared --iters 3 2>&1 | tail -24 ) > $O/h_trace_prepared_tpb128_r$r.log
done
( ZC_MSM_TRACE=2 timeout 120 python tools/run_msm.py --rank 0 --nranks 8 --fixed-base --iters 3 2>&1 | tail -14 ) > $O/h_trace_fb_r0.log
( ZC_MSM_ACC_TPB=128 ZC_MSM_TRACE=2 timeout 120 python tools/run_msm.py --rank 0 --nranks 8 --fixed-base --iters 3 2>&1 | tail -14 ) > $O/h_trace_fb_tpb128_r0.log
( timeout 120 python tools/run_msm.py --prepared --iters 5 2>&1 | tail -3 ) > $O/h_1gpu.log
( ZC_MSM_ACC_TPB=64 timeout 120 python tools/run_msm.py --prepared --iters 5 2>&1 | tail -3 ) > $O/h_1gpu_tpb64.log
( timeout 120 python tools/run_msm.py --rank 1 --nranks 4 --prepared --iters 3 2>&1 | tail -2 ) > $O/h_r1of4.log
( timeout 120 python tools/run_msm.py --rank 1 --nranks 2 --prepared --iters 3 2>&1 | tail -2 ) > $O/h_r1of2.log
cat $O/h_pytest.log; for f in h_trace_prepared_r7 h_trace_prepared_tpb128_r7 h_trace_prepared_r3 h_trace_prepared_tpb128_r3 h_trace_prepared_r0 h_trace_fb_r0 h_trace_fb_tpb128_r0 h_1gpu h_1gpu_tpb64 h_r1of4 h_r1of2; do echo "$f: $(tail -n 1 $O/$f.log)"; done
