#!/bin/bash
# 1 GPU: exact warp_sample with texel pairs (voxeliser, inject) — parity; pile-up test; sanitizer on the huge-list kernels; config 4 bench + ncu of its trace
TAG=${1:-r02t}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_modes.py -m gpu -q -k "not config5 and not config2" -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/${TAG}_pytest.txt
MODES="huge" bash tools/gpu_sanitize.sh ${TAG} > /dev/null 2>&1
timeout 400 python bench.py --config 4 --steps 50 --warmup 5 --no-cpu-baseline 2> gpurun_out/${TAG}_bench_c4_n1.err | tail -1 > gpurun_out/${TAG}_bench_c4_n1.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_cone_trace -s 2 -c 1 -o gpurun_out/${TAG}_trace_c4 -f python bench.py --config 4 --steps 2 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/${TAG}_ncu.log 2>&1
tail -6 gpurun_out/${TAG}_pytest.txt | cut -c1-300; grep -v "^=========  *$" gpurun_out/${TAG}_sanitizer.txt | cut -c1-200 | tail -12
for f in c4_n1; do python -c "
import json; j=json.loads(open('gpurun_out/${TAG}_bench_$f.json').read()); k=j['kernels_ms']; print('$f value', j['value'], 'e2e', j['e2e']['value'], k)" || tail -20 gpurun_out/${TAG}_bench_$f.err; done
