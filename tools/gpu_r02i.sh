#!/bin/bash
TAG=${1:-r02i}
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_sparse.py -m gpu -q -s -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/${TAG}_pytest_sharded.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tools/sharded_parity.py animated 2>&1 | grep "^{" | tail -1 > gpurun_out/${TAG}_sharded_parity_animated_n$N.txt
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus $N --config 3 --steps 100 --warmup 10 2> gpurun_out/${TAG}_bench_c3_n$N.err | tail -1 > gpurun_out/${TAG}_bench_c3_n$N.json
tail -12 gpurun_out/${TAG}_pytest_sharded.txt | cut -c1-300; cat gpurun_out/${TAG}_sharded_parity_animated_n$N.txt | cut -c1-500
python -c "
import json; j=json.loads(open('gpurun_out/${TAG}_bench_c3_n$N.json').read()); print('value', j['value'], 'e2e', j['e2e']['value'], j['kernels_ms'])" || tail -20 gpurun_out/${TAG}_bench_c3_n$N.err
