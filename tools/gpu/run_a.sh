#!/bin/bash
# round-2 GPU batch A: baseline tests, latency probes, pipe micro-benchmarks, per-rank MSM timelines, host topology
mkdir -p gpurun_out
O=gpurun_out
( timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) > $O/a_pytest.log
( cd tools/ubench && timeout 60 ./chainbench ) > $O/a_chainbench.log 2>&1
for b in pipes pipes2 pipes3 pipes4 latbench latbench2 latbench3 montbench karabench; do
  ( cd tools/ubench && echo "== $b" && timeout 120 ./$b ) >> $O/a_ubench.log 2>&1
done
for r in 7 0; do
  ( ZC_MSM_TRACE=2 timeout 120 python tools/run_msm.py --rank $r --nranks 8 --prepared --iters 3 2>&1 | tail -60 ) > $O/a_trace_prepared_r$r.log
done
( ZC_MSM_TRACE=2 timeout 120 python tools/run_msm.py --rank 7 --nranks 8 --iters 3 2>&1 | tail -60 ) > $O/a_trace_plain_r7.log
( ZC_MSM_TRACE=2 timeout 120 python tools/run_msm.py --prepared --iters 3 2>&1 | tail -120 ) > $O/a_trace_prepared_1gpu.log
( lscpu | head -30; echo; ls /sys/devices/system/node/; cat /sys/devices/system/node/node*/cpulist; nvidia-smi topo -m; for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/class 2>/dev/null)" = "0x030200" ]; then echo $d $(cat $d/numa_node) $(cat $d/local_cpulist); fi; done; nvidia-smi --query-gpu=pci.bus_id,clocks.sm,clocks.max.sm,power.limit --format=csv; free -g | head -3 ) > $O/a_topology.log 2>&1
tail -3 $O/a_pytest.log; cat $O/a_chainbench.log
