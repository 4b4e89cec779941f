#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
( CUDA_DEVICE_MAX_CONNECTIONS=32 ZC_PEER_TIMEOUT_MS=1500 timeout 120 python tools/gpu/probe_local.py 2 4 64 2>&1 | tail -12 ) > $O/g_probe_graph.log
( timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > $O/g_pytest.log
for r in 7 4 3 0; do
  ( ZC_MSM_TRACE=2 timeout 120 python tools/run_msm.py --rank $r --nranks 8 --prepared --iters 3 2>&1 | tail -30 ) > $O/g_trace_prepared_r$r.log
done
( ZC_MSM_TRACE=2 timeout 120 python tools/run_msm.py --rank 7 --nranks 8 --fixed-base --iters 3 2>&1 | tail -16 ) > $O/g_trace_fb_r7.log
( ZC_MSM_TRACE=2 timeout 120 python tools/run_msm.py --rank 0 --nranks 8 --fixed-base --iters 3 2>&1 | tail -16 ) > $O/g_trace_fb_r0.log
for mb in 4 8 32 64; do
  ( ZC_PIPE_CHUNK_MB=$mb timeout 300 python bench.py --steps 5 --warmup 3 --skip-extra --skip-cpu --no-sustain 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('chunk_mb $mb', d['e2e'])" ) >> $O/g_chunks.log 2>&1
done
cat $O/g_probe_graph.log; cat $O/g_pytest.log; for r in 7 4 3 0; do tail -n 1 $O/g_trace_prepared_r$r.log; done; tail -n 1 $O/g_trace_fb_r7.log $O/g_trace_fb_r0.log; cat $O/g_chunks.log
