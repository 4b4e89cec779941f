#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2_surface.py -m gpu -x -q -k "invert or affine or div or compress or decompress or elligator or vector_ops or sqrt" 2>&1 | tail -4 ) > $O/ai.log
( timeout 600 python bench.py --steps 5 --warmup 3 --skip-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); f=d['section8f_next_rows']; print({k: round(f[k],3) for k in f if k.endswith('_ms')}); print(f['roofline_int_fe_invert']['frac'])" ) >> $O/ai.log 2>&1
cat $O/ai.log
