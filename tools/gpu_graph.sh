#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 | tee gpurun_out/pytest_gpu.txt
timeout 600 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/bench_graph.json 2> gpurun_out/bench_graph.err; tail -3 gpurun_out/bench_graph.err
timeout 600 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-graph > gpurun_out/bench_nograph.json 2> gpurun_out/bench_nograph.err
python - <<'PY'
import json
for f in ("bench_graph","bench_nograph"):
    try:
        j=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f,"value",j["value"],"e2e",j["e2e"]["value"],"launches",j["gpu_launches"],j["config"]["cuda_graph"])
    except Exception as e: print(f,"parse failed",e)
PY
