#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
: > $O/af.log
for r in 1 0 1 0; do
  ( ZC_PIPE_RAMP=$r timeout 300 python bench.py --steps 10 --warmup 3 --skip-extra --skip-cpu --no-sustain 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ramp $r', d['e2e']['ms_per_step'], d['e2e']['matches_resident_path'], d['packed_wire_format']['e2e']['ms_per_step'], d['packed_wire_format']['matches_limb_layout_results'])" ) >> $O/af.log 2>&1
done
( timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -k "field or scalar_ops or point_ops or cfg2 or cfg3 or empty" 2>&1 | tail -3 ) >> $O/af.log
cat $O/af.log
