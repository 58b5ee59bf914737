#!/bin/bash
# 1 GPU: multisample voxelisation parity + no regression of the default path (parity files, config 3 bench)
TAG=${1:-r02w}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_modes.py tests/test_gpu_parity.py tests/test_gpu_sparse.py -m gpu -q -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/${TAG}_pytest.txt
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline 2> gpurun_out/${TAG}_bench_c3_n1.err | tail -1 > gpurun_out/${TAG}_bench_c3_n1.json
tail -8 gpurun_out/${TAG}_pytest.txt | cut -c1-400
for f in c3_n1; do python -c "
import json; j=json.loads(open('gpurun_out/${TAG}_bench_$f.json').read()); k=j['kernels_ms']; print('$f value', j['value'], 'e2e', j['e2e']['value'], k)" || tail -20 gpurun_out/${TAG}_bench_$f.err; done
