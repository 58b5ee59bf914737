#!/bin/bash
# 1 GPU: cone-trace variant check — config 4 parity test + config 4 and config 3 bench lines
TAG=${1:-r02o}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -q -k "warped_frame or config4 or final_image or sponza_256" -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/${TAG}_pytest.txt
timeout 400 python bench.py --config 4 --steps 50 --warmup 5 --no-cpu-baseline 2> gpurun_out/${TAG}_bench_c4_n1.err | tail -1 > gpurun_out/${TAG}_bench_c4_n1.json
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline 2> gpurun_out/${TAG}_bench_c3_n1.err | tail -1 > gpurun_out/${TAG}_bench_c3_n1.json
tail -6 gpurun_out/${TAG}_pytest.txt | cut -c1-300
for f in c4_n1 c3_n1; do python -c "
import json; j=json.loads(open('gpurun_out/${TAG}_bench_$f.json').read()); k=j['kernels_ms']; print('$f value', j['value'], 'e2e', j['e2e']['value'], 'trace', k['k_cone_trace'])" || tail -20 gpurun_out/${TAG}_bench_$f.err; done
