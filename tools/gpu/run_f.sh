#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32 ZC_PEER_TIMEOUT_MS=1500
( timeout 120 python tools/gpu/probe_local.py 2 4 64 2>&1 | tail -20 ) > $O/f_probe_graph.log
( ZC_MSM_TRACE=1 timeout 120 python tools/gpu/probe_local.py 2 4 64 2>&1 | grep -v "us  s\|zc_msm trace" | tail -20 ) > $O/f_probe_nograph.log
( timeout 120 python tools/gpu/probe_local.py 2 4 5000 2>&1 | tail -20 ) > $O/f_probe_graph_5000.log
cat $O/f_probe_graph.log; echo; cat $O/f_probe_nograph.log; echo; cat $O/f_probe_graph_5000.log
