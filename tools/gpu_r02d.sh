#!/bin/bash
# round 2, fourth GPU call: config 4 (pile-up voxel) pass by pass, all GPU tests but the 8 Mi soup, bench config 3 and 4
TAG=r02d
mkdir -p gpurun_out
timeout 120 python tools/diag_config4.py 128 480 270 2 > gpurun_out/${TAG}_diag_c4_128.txt 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_diag_c4_128.txt
timeout 200 python tools/diag_config4.py 512 480 270 2 > gpurun_out/${TAG}_diag_c4_512.txt 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_diag_c4_512.txt
timeout 1200 python -m pytest tests -m gpu -q -s -k "not capacity" -p no:cacheprovider 2>&1 | tail -70 > gpurun_out/${TAG}_pytest_gpu.txt
timeout 300 python bench.py --steps 200 --warmup 20 2> gpurun_out/${TAG}_bench.err | tail -1 > gpurun_out/${TAG}_bench.json
timeout 300 python bench.py --config 4 --steps 30 --warmup 5 --no-cpu-baseline 2> gpurun_out/${TAG}_bench_c4.err | tail -1 > gpurun_out/${TAG}_bench_c4.json
tail -14 gpurun_out/${TAG}_diag_c4_128.txt; tail -30 gpurun_out/${TAG}_diag_c4_512.txt; tail -25 gpurun_out/${TAG}_pytest_gpu.txt | cut -c1-300
for f in bench bench_c4; do python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/${TAG}_$f.json").read())
    print("$f", j["value"], "e2e", j["e2e"]["value"], "launches", j.get("gpu_launches"), j["kernels_ms"], j["roofline_hbm"]["frac"], j.get("passes_ms"))
except Exception as e:
    print("$f parse failed", e); print(open("gpurun_out/${TAG}_$f.err").read()[-1500:])
PY
done
