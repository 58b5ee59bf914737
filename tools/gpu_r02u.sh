#!/bin/bash
# 1 GPU: chunked list walk for voxels beyond 1024 fragments without a huge-table slot — long-list tests, config 4 parity, sanitizer, config 4 bench
TAG=${1:-r02u}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_configs.py tests/test_gpu_modes.py -m gpu -q -k "config4 or long_per_voxel or pile_up" -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/${TAG}_pytest.txt
MODES="long huge" bash tools/gpu_sanitize.sh ${TAG} > /dev/null 2>&1
timeout 400 python bench.py --config 4 --steps 50 --warmup 5 --no-cpu-baseline 2> gpurun_out/${TAG}_bench_c4_n1.err | tail -1 > gpurun_out/${TAG}_bench_c4_n1.json
tail -6 gpurun_out/${TAG}_pytest.txt | cut -c1-300; grep -v "^=========  *$" gpurun_out/${TAG}_sanitizer.txt | cut -c1-200 | tail -24
for f in c4_n1; do python -c "
import json; j=json.loads(open('gpurun_out/${TAG}_bench_$f.json').read()); k=j['kernels_ms']; print('$f value', j['value'], 'e2e', j['e2e']['value'], {a:b for a,b in k.items() if 'voxel' in a})" || tail -20 gpurun_out/${TAG}_bench_$f.err; done
