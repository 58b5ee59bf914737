#!/bin/bash
TAG=${1:-r02k}
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_configs.py tests/test_gpu_modes.py -m gpu -q -s -k "not capacity" -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/${TAG}_pytest.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tools/sharded_parity.py animated 2>&1 | grep "^{" | tail -1 > gpurun_out/${TAG}_sharded_parity_animated_n$N.txt
timeout 400 python bench.py --config 4 --steps 50 --warmup 5 --no-cpu-baseline 2> gpurun_out/${TAG}_bench_c4_n1.err | tail -1 > gpurun_out/${TAG}_bench_c4_n1.json
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29624 bench.py --gpus $N --config 4 --steps 50 --warmup 5 2> gpurun_out/${TAG}_bench_c4_n$N.err | tail -1 > gpurun_out/${TAG}_bench_c4_n$N.json
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline 2> gpurun_out/${TAG}_bench_c3_n1.err | tail -1 > gpurun_out/${TAG}_bench_c3_n1.json
tail -12 gpurun_out/${TAG}_pytest.txt | cut -c1-300; cat gpurun_out/${TAG}_sharded_parity_animated_n$N.txt | cut -c1-500
for f in c4_n1 c4_n$N c3_n1; do python -c "
import json; j=json.loads(open('gpurun_out/${TAG}_bench_$f.json').read()); print('$f value', j['value'], 'e2e', j['e2e']['value'], j['kernels_ms'], j.get('passes_ms'))" || tail -20 gpurun_out/${TAG}_bench_$f.err; done
