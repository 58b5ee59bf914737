#!/bin/bash
# Round-end check on one B200: GPU tests (through the C ABI), smoke, default bench line + reference arm.   usage: gpu_final.sh [tag]
TAG=${1:-final}
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -v "^$" | tail -25 | tee gpurun_out/${TAG}_pytest_gpu.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/${TAG}_smoke.txt
timeout 300 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -3 gpurun_out/${TAG}_bench.err
if [ "${REF_ARM:-0}" = "1" ]; then
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
tail -2 gpurun_out/${TAG}_bench_reference.err; cut -c1-400 gpurun_out/${TAG}_bench_reference.json
fi
python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
    print("value",j["value"],"e2e",j["e2e"]["value"],"launches",j["gpu_launches"],"clocks",j["clocks"])
    print("voxel_passes",j["voxel_passes"]); print("cpu",j["cpu_baseline"])
except Exception as e: print("bench parse failed",e)
PY
