#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4 --steps 10 --warmup 3 > $O/4_bench.json 2> $O/4_bench.err )
tail -3 $O/4_bench.err | cut -c1-300; python - <<'P'
import json
try:
    d = json.loads(open("gpurun_out/4_bench.json").read().strip().splitlines()[-1])
    print(d["value"], d["e2e"]["value"], json.dumps(d.get("msm_scaling"))[:900])
    print("oracle", d["config5_msm"].get("matches_oracle"), d["config5_msm"].get("ms_per_msm_nccl_exchange"))
except Exception as e:
    print("ERR", e)
P
