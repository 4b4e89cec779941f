#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 ) > $O/san5.log
( timeout 600 compute-sanitizer --tool initcheck --print-limit 40 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "msm_vs_naive and 1000" 2>&1 | grep -E "^=========     at|ERROR SUMMARY|passed" | sed "s/+0x[0-9a-f]*//" | cut -c1-160 | sort | uniq -c | sort -rn | head -12 ) >> $O/san5.log
for mode in --prepared "--fixed-base --rank 3 --nranks 8"; do ( echo -n "[$mode] "; timeout 120 python tools/run_msm.py $mode --iters 8 2>&1 | grep "msm n=" | tail -6 | awk '{print $6}' | sort -n | head -1 ) >> $O/san5.log; done
cat $O/san5.log
