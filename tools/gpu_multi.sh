#!/bin/bash
# multi-GPU round: gpu_multi.sh N   (parity vs single GPU, then the bench at N for config 3 and the 512^3/4K config)
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for w in room sponza sponza512; do
  timeout 600 $TR tools/sharded_parity.py $w > gpurun_out/parity_$w.log 2>&1
  grep -E '^\{' gpurun_out/parity_$w.log | tail -1 | tee -a gpurun_out/sharded_parity_n$N.txt
  grep -E '^\{' gpurun_out/parity_$w.log > /dev/null || tail -25 gpurun_out/parity_$w.log
done
timeout 600 $TR bench.py --gpus $N --steps 100 --warmup 10 2> gpurun_out/bench_n$N.err | tail -1 > gpurun_out/bench_n$N.json
timeout 600 $TR bench.py --gpus $N --steps 50 --warmup 5 --dim 512 --width 3840 --height 2160 2> gpurun_out/bench512_n$N.err | tail -1 > gpurun_out/bench512_n$N.json
VCT_SPARSE_EXCHANGE=0 timeout 600 $TR bench.py --gpus $N --steps 50 --warmup 5 --dim 512 --width 3840 --height 2160 2> gpurun_out/bench512_dense_n$N.err | tail -1 > gpurun_out/bench512_dense_n$N.json
if [ -z "$SKIP_N1" ]; then timeout 600 python bench.py --steps 50 --warmup 5 --dim 512 --width 3840 --height 2160 --no-cpu-baseline 2> gpurun_out/bench512_n1.err | tail -1 > gpurun_out/bench512_n1.json; fi
python - <<PY
import json
for f in ("bench_n$N","bench512_n$N","bench512_dense_n$N","bench512_n1"):
    try:
        j=json.loads(open(f"gpurun_out/{f}.json").read())
        print(f, "value",j["value"],"e2e",j["e2e"]["value"],"launches",j["gpu_launches"], {k:v for k,v in list(j["kernels_ms"].items())[:9]})
    except Exception as e: print(f,"parse failed",e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
PY
