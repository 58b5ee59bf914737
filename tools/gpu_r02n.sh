#!/bin/bash
# 1 GPU: the 32-CTA warp map kernel — parity (warp map volumes, warped frames, config 4), config 4 bench, and one ncu --set full capture of
# config 4's cone trace (WARP_TEXTURE instantiation)
TAG=${1:-r02n}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -q -k "warp or variants or config4 or cpp or glsl" -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/${TAG}_pytest.txt
timeout 400 python bench.py --config 4 --steps 50 --warmup 5 --no-cpu-baseline 2> gpurun_out/${TAG}_bench_c4_n1.err | tail -1 > gpurun_out/${TAG}_bench_c4_n1.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_cone_trace -s 2 -c 1 -o gpurun_out/${TAG}_trace_c4 -f python bench.py --config 4 --steps 2 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/${TAG}_ncu.log 2>&1
tail -6 gpurun_out/${TAG}_pytest.txt | cut -c1-300
for f in c4_n1; do python -c "
import json; j=json.loads(open('gpurun_out/${TAG}_bench_$f.json').read()); print('$f value', j['value'], 'e2e', j['e2e']['value'], j['kernels_ms'])" || tail -20 gpurun_out/${TAG}_bench_$f.err; done
tail -3 gpurun_out/${TAG}_ncu.log; ls -la gpurun_out/${TAG}_trace_c4.ncu-rep
