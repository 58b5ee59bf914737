#!/bin/bash
# 1 GPU: warp map pair table (256-bit loads) in the cone tracer — parity (module order kept), config 4 / 3 bench
TAG=${1:-r02p}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -q -k "not config5 and not config2" -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/${TAG}_pytest.txt
timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "final_image" -p no:cacheprovider 2>&1 | tail -5 > gpurun_out/${TAG}_pytest_trace_first.txt
timeout 400 python bench.py --config 4 --steps 50 --warmup 5 --no-cpu-baseline 2> gpurun_out/${TAG}_bench_c4_n1.err | tail -1 > gpurun_out/${TAG}_bench_c4_n1.json
tail -6 gpurun_out/${TAG}_pytest.txt | cut -c1-300; tail -3 gpurun_out/${TAG}_pytest_trace_first.txt | cut -c1-300
for f in c4_n1; do python -c "
import json; j=json.loads(open('gpurun_out/${TAG}_bench_$f.json').read()); k=j['kernels_ms']; print('$f value', j['value'], 'e2e', j['e2e']['value'], 'trace', k['k_cone_trace'])" || tail -20 gpurun_out/${TAG}_bench_$f.err; done
