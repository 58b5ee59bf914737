#!/bin/bash
# quick check of new GPU tests + a short bench line.   usage: gpu_quick2.sh <tag> <pytest args...>
TAG=$1; shift
mkdir -p gpurun_out
timeout 300 python -m pytest "$@" -m gpu -x -q -s 2>&1 | tail -40 | tee gpurun_out/${TAG}_pytest.txt
timeout 200 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
    print("value",j["value"],"e2e",j["e2e"]["value"],"kernels",j["kernels_ms"])
except Exception as e: print("bench parse failed",e); print(open('gpurun_out/${TAG}_bench.err').read()[-1500:])
PY
