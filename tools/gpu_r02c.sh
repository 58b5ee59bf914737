#!/bin/bash
# round 2, third GPU call: config-4 diagnostic (bounded), GPU tests without the heavy config cases, bench of the new voxeliser
TAG=r02c
mkdir -p gpurun_out
timeout 120 python tools/diag_config4.py 128 480 270 2 > gpurun_out/${TAG}_diag_c4_128.txt 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_diag_c4_128.txt
timeout 200 python tools/diag_config4.py 512 480 270 2 > gpurun_out/${TAG}_diag_c4_512.txt 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_diag_c4_512.txt
timeout 900 python -m pytest tests -m gpu -q -s -k "not config4 and not capacity" -p no:cacheprovider 2>&1 | tail -70 > gpurun_out/${TAG}_pytest_gpu.txt
timeout 300 python bench.py --steps 200 --warmup 20 2> gpurun_out/${TAG}_bench.err | tail -1 > gpurun_out/${TAG}_bench.json
tail -12 gpurun_out/${TAG}_diag_c4_128.txt; tail -16 gpurun_out/${TAG}_diag_c4_512.txt; tail -25 gpurun_out/${TAG}_pytest_gpu.txt | cut -c1-300
python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/${TAG}_bench.json").read())
    print("bench", j["value"], "e2e", j["e2e"]["value"], "launches", j.get("gpu_launches"), j["kernels_ms"], j["roofline_hbm"]["frac"])
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/${TAG}_bench.err").read()[-1500:])
PY
