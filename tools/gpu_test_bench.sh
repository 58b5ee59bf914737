#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 | tee gpurun_out/pytest_gpu.txt
timeout 600 python bench.py --steps ${STEPS:-50} --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -5 gpurun_out/bench.err
python - <<'PY'
import json
try:
    j=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
    print("value",j["value"],"e2e",j["e2e"]["value"],"launches",j["gpu_launches"],"clocks",j["clocks"])
    print("kernels_ms",j["kernels_ms"])
    print("passes_ms",j["passes_ms"])
    print("voxel_passes",j["voxel_passes"])
    for r in j["roofline_passes"]: print(r["kernel"],r["ms"],r["achieved"],r["frac"])
    print("cpu",j["cpu_baseline"])
except Exception as e: print("bench parse failed",e)
PY
