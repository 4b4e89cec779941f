#!/bin/bash
# two GPUs: the whole GPU test suite (the two-GPU test included) + bench.py at N = 2 and N = 1
mkdir -p gpurun_out
O=gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > $O/j_pytest.log
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > $O/j_bench_n2.json 2> $O/j_bench_n2.err )
( timeout 600 python bench.py --steps 10 --warmup 3 > $O/j_bench_n1.json 2> $O/j_bench_n1.err )
cat $O/j_pytest.log; tail -5 $O/j_bench_n2.err | cut -c1-300; python - <<'P'
import json
for f in ("gpurun_out/j_bench_n2.json", "gpurun_out/j_bench_n1.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["e2e"]["value"], json.dumps(d.get("msm_scaling"))[:700])
        print(" packed", json.dumps(d.get("packed_wire_format"))[:600])
        print(" oracle", d["config5_msm"].get("matches_oracle"), d.get("host_placement"), d["roofline_int"]["frac"])
    except Exception as e:
        print(f, "ERR", e)
P
