#!/bin/bash
# N GPUs (8): interleaved stripes against contiguous slabs — parity once, then config 3 / 4 bench lines
TAG=${1:-r02q}
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
VCT_SLAB_STRIPE=16 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tools/sharded_parity.py animated 2>&1 | grep "^{" | tail -1 > gpurun_out/${TAG}_parity_animated_stripe16_n$N.txt
for CS in "3 0" "3 16" "3 32" "4 0" "4 16"; do set -- $CS; C=$1; S=$2
VCT_SLAB_STRIPE=$S timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29624 bench.py --gpus $N --config $C --steps $([ $C = 3 ] && echo 200 || echo 50) --warmup 10 2> gpurun_out/${TAG}_bench_c${C}_n${N}_stripe$S.err | tail -1 > gpurun_out/${TAG}_bench_c${C}_n${N}_stripe$S.json
done
cat gpurun_out/${TAG}_parity_*_n$N.txt | cut -c1-700
for f in c3_n${N}_stripe0 c3_n${N}_stripe16 c3_n${N}_stripe32 c4_n${N}_stripe0 c4_n${N}_stripe16; do python -c "
import json; j=json.loads(open('gpurun_out/${TAG}_bench_$f.json').read()); print('$f value', j['value'], 'e2e', j['e2e']['value'], j['kernels_ms'])" || tail -20 gpurun_out/${TAG}_bench_$f.err; done
