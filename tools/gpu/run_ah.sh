#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2_surface.py tests/test_cpp_surface.py -m gpu -x -q -k "invert or affine or div or surface or empty or operator" 2>&1 | tail -4 ) > $O/ah.log
( timeout 600 python bench.py --steps 5 --warmup 3 --skip-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); f=d['section8f_next_rows']; print({k: f[k] for k in f if k.endswith('_ms')})" ) >> $O/ah.log 2>&1
cat $O/ah.log
