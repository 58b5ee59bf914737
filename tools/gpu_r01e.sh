#!/bin/bash
# round r01e: GPU tests, frame breakdown sparse vs dense, bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.txt
echo "--- sparse"; python tools/profile_frame.py 4 --kernels 2>&1 | tail -32 | tee gpurun_out/kernels_frame.txt
echo "--- dense";  VCT_SPARSE=0 python tools/profile_frame.py 4 --kernels 2>&1 | tail -32 | tee gpurun_out/kernels_frame_dense.txt
timeout 600 python bench.py --steps ${STEPS:-200} --warmup 20 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -5 gpurun_out/bench.err
python - <<'PY'
import json
try:
    j=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
    print("value",j["value"],"e2e",j["e2e"]["value"],"launches",j["gpu_launches"],"clocks",j["clocks"])
    print("kernels_ms",j["kernels_ms"])
    print("voxel_passes",j["voxel_passes"])
    for r in j["roofline_passes"]: print(r["kernel"],r["ms"],r["achieved"],r["frac"])
except Exception as e: print("bench parse failed",e)
PY
