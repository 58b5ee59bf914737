#!/bin/bash
# N-GPU confirmation (N = 4 or 8): parity of the Sponza frame vs one GPU, then the bench at N for config 3 and 512^3/4K
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 400 $TR tools/sharded_parity.py sponza > gpurun_out/parity_sponza_n$N.log 2>&1
grep -E '^\{' gpurun_out/parity_sponza_n$N.log | tail -1 | tee gpurun_out/sharded_parity_n$N.txt
grep -E '^\{' gpurun_out/parity_sponza_n$N.log > /dev/null || tail -25 gpurun_out/parity_sponza_n$N.log
timeout 400 $TR bench.py --gpus $N --steps 200 --warmup 20 2> gpurun_out/bench_n$N.err | tail -1 > gpurun_out/bench_n$N.json
timeout 400 $TR bench.py --gpus $N --steps 100 --warmup 10 --dim 512 --width 3840 --height 2160 2> gpurun_out/bench512_n$N.err | tail -1 > gpurun_out/bench512_n$N.json
python - <<PY
import json
for f in ("bench_n$N","bench512_n$N"):
    try:
        j=json.loads(open(f"gpurun_out/{f}.json").read())
        print(f, "value",j["value"],"e2e",j["e2e"]["value"],"launches",j["gpu_launches"], {k:v for k,v in list(j["kernels_ms"].items())[:12]})
    except Exception as e: print(f,"parse failed",e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
PY
