#!/bin/bash
# 2 (or N) GPUs: interleaved stripes (VCT_SLAB_STRIPE=16) against contiguous slabs — parity tests, then config 3 / 4 bench lines both ways
TAG=${1:-ab}
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_modes.py -m gpu -q -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/${TAG}_pytest_contiguous.txt
VCT_SLAB_STRIPE=16 timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -q -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/${TAG}_pytest_stripe16.txt
for W in sponza animated; do
VCT_SLAB_STRIPE=16 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tools/sharded_parity.py $W 2>&1 | grep "^{" | tail -1 > gpurun_out/${TAG}_parity_${W}_stripe16_n$N.txt
done
for S in 0 16; do for C in 3 4; do
VCT_SLAB_STRIPE=$S timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29624 bench.py --gpus $N --config $C --steps $([ $C = 3 ] && echo 200 || echo 50) --warmup 10 2> gpurun_out/${TAG}_bench_c${C}_n${N}_stripe$S.err | tail -1 > gpurun_out/${TAG}_bench_c${C}_n${N}_stripe$S.json
done; done
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline 2> gpurun_out/${TAG}_bench_c3_n1.err | tail -1 > gpurun_out/${TAG}_bench_c3_n1.json
tail -6 gpurun_out/${TAG}_pytest_contiguous.txt | cut -c1-300; tail -6 gpurun_out/${TAG}_pytest_stripe16.txt | cut -c1-300; cat gpurun_out/${TAG}_parity_*_n$N.txt | cut -c1-600
for f in c3_n${N}_stripe0 c3_n${N}_stripe16 c4_n${N}_stripe0 c4_n${N}_stripe16 c3_n1; do python -c "
import json; j=json.loads(open('gpurun_out/${TAG}_bench_$f.json').read()); print('$f value', j['value'], 'e2e', j['e2e']['value'], j['kernels_ms'])" || tail -20 gpurun_out/${TAG}_bench_$f.err; done
