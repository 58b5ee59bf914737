#!/bin/bash
# 1 GPU: adaptive small-item chunk size in k_voxel_tiles — parity + bench of configs 2, 3, 5, 4
TAG=${1:-r02ac}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_modes.py tests/test_gpu_configs.py -m gpu -q -k "not capacity" -p no:cacheprovider 2>&1 | tail -4 > gpurun_out/${TAG}_pytest.txt
for C in 2 3 5 4; do
timeout 400 python bench.py --config $C --steps $([ $C -le 3 ] && echo 200 || echo 50) --warmup 10 --no-cpu-baseline 2> gpurun_out/${TAG}_bench_c${C}_n1.err | tail -1 > gpurun_out/${TAG}_bench_c${C}_n1.json
done
tail -3 gpurun_out/${TAG}_pytest.txt | cut -c1-300
for f in c2_n1 c3_n1 c5_n1 c4_n1; do python -c "
import json; j=json.loads(open('gpurun_out/${TAG}_bench_$f.json').read()); k=j['kernels_ms']; print('$f value', j['value'], 'e2e', j['e2e']['value'], {a:b for a,b in k.items() if 'voxel_tiles' in a or 'voxel_bin' in a})" || tail -20 gpurun_out/${TAG}_bench_$f.err; done
