#!/bin/bash
TAG=r02f
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_modes.py tests/test_gpu_configs.py tests/test_gpu_sparse.py tests/test_gpu_parity.py -m gpu -q -s -k "not capacity" -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/${TAG}_pytest_gpu.txt
timeout 300 python bench.py --steps 200 --warmup 20 2> gpurun_out/${TAG}_bench.err | tail -1 > gpurun_out/${TAG}_bench.json
timeout 300 python bench.py --config 4 --steps 50 --warmup 5 --no-cpu-baseline 2> gpurun_out/${TAG}_bench_c4.err | tail -1 > gpurun_out/${TAG}_bench_c4.json
tail -12 gpurun_out/${TAG}_pytest_gpu.txt | cut -c1-300
for f in bench bench_c4; do python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/${TAG}_$f.json").read())
    print("$f", j["value"], "e2e", j["e2e"]["value"], "launches", j.get("gpu_launches"), j["kernels_ms"], j["roofline_hbm"]["frac"], j.get("passes_ms"))
except Exception as e:
    print("$f parse failed", e); print(open("gpurun_out/${TAG}_$f.err").read()[-1500:])
PY
done
