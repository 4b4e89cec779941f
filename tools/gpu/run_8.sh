#!/bin/bash
# eight GPUs: the sharded MSM tool (seconds) and bench.py at N = 8 as the driver launches it
mkdir -p gpurun_out
O=gpurun_out
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tools/run_msm_sharded.py --iters 20 > $O/8_msm.log 2>&1 )
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --steps 10 --warmup 3 > $O/8_bench.json 2> $O/8_bench.err )
tail -2 $O/8_msm.log | cut -c1-700; tail -3 $O/8_bench.err | cut -c1-300; python - <<'P'
import json
try:
    d = json.loads(open("gpurun_out/8_bench.json").read().strip().splitlines()[-1])
    print(d["value"], d["e2e"]["value"], json.dumps(d.get("msm_scaling"))[:900])
    print("oracle", d["config5_msm"].get("matches_oracle"), d.get("host_placement"), d["config5_msm"].get("ms_per_msm_nccl_exchange"), d["config5_msm"].get("ms_per_msm_incl_scalar_broadcast"))
except Exception as e:
    print("ERR", e)
P
