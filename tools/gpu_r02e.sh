#!/bin/bash
# round 2, multi-GPU call (gpurun --gpus N): sharded parity tests (torchrun + single-process handle + C++ host), bench at 1..N GPUs
TAG=${1:-r02e}
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -q -s -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/${TAG}_pytest_sharded.txt
for w in sponza animated; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tools/sharded_parity.py $w 2>&1 | grep "^{" | tail -1 > gpurun_out/${TAG}_sharded_parity_${w}_n$N.txt
done
for g in 1 2 4 8; do
  [ $g -le $N ] || continue
  for c in 3 4; do
    if [ $g -eq 1 ]; then timeout 400 python bench.py --config $c --steps 100 --warmup 10 --no-cpu-baseline 2> gpurun_out/${TAG}_bench_c${c}_n$g.err | tail -1 > gpurun_out/${TAG}_bench_c${c}_n$g.json
    else timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 2962$g bench.py --gpus $g --config $c --steps 100 --warmup 10 2> gpurun_out/${TAG}_bench_c${c}_n$g.err | tail -1 > gpurun_out/${TAG}_bench_c${c}_n$g.json; fi
  done
done
tail -15 gpurun_out/${TAG}_pytest_sharded.txt | cut -c1-400; cat gpurun_out/${TAG}_sharded_parity_*.txt | cut -c1-600
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench_c*_n*.json")):
    try:
        j = json.loads(open(f).read())
        print(f.split("/")[-1], "value", j["value"], "e2e", j["e2e"]["value"], "launches", j.get("gpu_launches"), j["execution"]["cuda_graph"][:40], {k: v for k, v in list(j["kernels_ms"].items())[:8]})
    except Exception as e:
        print(f, "parse failed", e); print(open(f.replace(".json", ".err")).read()[-1200:])
PY
