#!/bin/bash
# First GPU call of the next round: the GPU tests written after round 1's budget was spent, then the
# A/B of the split cone trace (VCT_TRACE_VARIANT bits 6-9: kernel times, whole-step time, PSNR against the default path), then the
# default bench line.   usage: gpurun --timeout 900 -- 'bash tools/gpu_next_round_first.sh [tag]'
TAG=${1:-r02a}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_tess_warp_gpu.py tests/test_zz_abi_validation_gpu.py -m gpu -q -s 2>&1 | tail -30 | tee gpurun_out/${TAG}_pytest_new.txt
timeout 400 python tools/trace_variants.py 20 0 64 192 320 448 576 704 2>&1 | tail -12 | tee gpurun_out/${TAG}_trace_split_ab.txt
timeout 300 python bench.py --steps 200 --warmup 20 2> gpurun_out/${TAG}_bench.err | tail -1 > gpurun_out/${TAG}_bench.json
for v in 192 448; do VCT_TRACE_VARIANT=$v timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline 2> gpurun_out/${TAG}_bench_v$v.err | tail -1 > gpurun_out/${TAG}_bench_v$v.json; done
python - <<PY
import json
for f in ("${TAG}_bench", "${TAG}_bench_v192", "${TAG}_bench_v448"):
    try:
        j = json.loads(open(f"gpurun_out/{f}.json").read())
        print(f, "value", j["value"], "e2e", j["e2e"]["value"], "launches", j["gpu_launches"], j["config"]["cuda_graph"][:60], {k: v for k, v in list(j["kernels_ms"].items())[:6]})
    except Exception as e:
        print(f, "parse failed", e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
PY
