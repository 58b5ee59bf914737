#!/bin/bash
# short N-GPU check: Sponza parity vs one GPU + the default bench at N
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 200 $TR tools/sharded_parity.py sponza > gpurun_out/parity_sponza_n$N.log 2>&1
grep -E '^\{' gpurun_out/parity_sponza_n$N.log | tail -1 | cut -c1-120; grep -o '"ok": [a-z]*' gpurun_out/parity_sponza_n$N.log | tail -1
timeout 200 $TR bench.py --gpus $N --steps 100 --warmup 10 2> gpurun_out/bench_n$N.err | tail -1 > gpurun_out/bench_n$N.json
python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/bench_n$N.json").read())
    print("bench_n$N value",j["value"],"e2e",j["e2e"]["value"],"launches",j["gpu_launches"],j["config"]["cuda_graph"], {k:v for k,v in list(j["kernels_ms"].items())[:10]})
except Exception as e: print("parse failed",e); print(open("gpurun_out/bench_n$N.err").read()[-1500:])
PY
