#!/bin/bash
# 1 GPU: parallel digit pick in k_voxel_huge_select + staged replay — config 4 parity and bench
TAG=${1:-r02m}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_configs.py tests/test_gpu_modes.py -m gpu -q -k "config4 or long_per_voxel or config1" -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/${TAG}_pytest.txt
timeout 400 python bench.py --config 4 --steps 50 --warmup 5 --no-cpu-baseline 2> gpurun_out/${TAG}_bench_c4_n1.err | tail -1 > gpurun_out/${TAG}_bench_c4_n1.json
tail -6 gpurun_out/${TAG}_pytest.txt | cut -c1-300
for f in c4_n1; do python -c "
import json; j=json.loads(open('gpurun_out/${TAG}_bench_$f.json').read()); print('$f value', j['value'], 'e2e', j['e2e']['value'], j['kernels_ms'])" || tail -20 gpurun_out/${TAG}_bench_$f.err; done
