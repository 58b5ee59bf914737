#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.txt
timeout 600 python tools/trace_variants.py 10 0 2>&1 | tail -12 | tee gpurun_out/trace_variants.txt
python tools/profile_frame.py 4 --kernels 2>&1 | tail -32 | tee gpurun_out/kernels_frame.txt
timeout 600 python bench.py --steps ${STEPS:-200} --warmup 20 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -5 gpurun_out/bench.err
python - <<'PY'
import json
try:
    j=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
    print("value",j["value"],"e2e",j["e2e"]["value"],"launches",j["gpu_launches"],"clocks",j["clocks"])
    print("kernels_ms",j["kernels_ms"])
    print("voxel_passes",j["voxel_passes"])
except Exception as e: print("bench parse failed",e)
PY
