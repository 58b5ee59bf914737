#!/bin/bash
# 8-GPU check: sharded parity at N = all GPUs, bench config 3 and 4 at N
TAG=${1:-r02j}
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
for w in sponza animated; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tools/sharded_parity.py $w 2>&1 | grep "^{" | tail -1 > gpurun_out/${TAG}_sharded_parity_${w}_n$N.txt
done
for c in 3 4; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2962$c bench.py --gpus $N --config $c --steps 100 --warmup 10 2> gpurun_out/${TAG}_bench_c${c}_n$N.err | tail -1 > gpurun_out/${TAG}_bench_c${c}_n$N.json
done
cat gpurun_out/${TAG}_sharded_parity_*_n$N.txt | cut -c1-500
for c in 3 4; do python -c "
import json; j=json.loads(open('gpurun_out/${TAG}_bench_c${c}_n$N.json').read()); print('config $c N=$N value', j['value'], 'e2e', j['e2e']['value'], j['execution']['cuda_graph'][:30], j['kernels_ms'], j.get('passes_ms'))" || tail -20 gpurun_out/${TAG}_bench_c${c}_n$N.err; done
