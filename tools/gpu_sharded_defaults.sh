#!/bin/bash
# N GPUs (4 or 8): the defaults (stripes of 16 layers from 4 ranks on, mip chain pushing its levels to the peers) — parity, then config 3 / 4 bench lines;
# config 3 also with contiguous slabs (VCT_SLAB_STRIPE=-1)
TAG=${1:-defaults}
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
for W in sponza animated; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tools/sharded_parity.py $W 2>&1 | grep "^{" | tail -1 > gpurun_out/${TAG}_parity_${W}_n$N.txt
done
for CS in "3 0" "4 0" "3 -1"; do set -- $CS; C=$1; S=$2
VCT_SLAB_STRIPE=$S timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29624 bench.py --gpus $N --config $C --steps $([ $C = 3 ] && echo 200 || echo 50) --warmup 10 2> gpurun_out/${TAG}_bench_c${C}_n${N}_stripe$S.err | tail -1 > gpurun_out/${TAG}_bench_c${C}_n${N}_stripe$S.json
done
cat gpurun_out/${TAG}_parity_*_n$N.txt | cut -c1-700
for f in c3_n${N}_stripe0 c4_n${N}_stripe0 c3_n${N}_stripe-1; do python -c "
import json; j=json.loads(open('gpurun_out/${TAG}_bench_$f.json').read()); print('$f value', j['value'], 'e2e', j['e2e']['value'], j['kernels_ms'])" || tail -20 gpurun_out/${TAG}_bench_$f.err; done
