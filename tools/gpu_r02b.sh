#!/bin/bash
# round 2, second GPU call: all GPU tests (new: modes, configs), bench lines of every configuration, soup capacity sweep
TAG=r02b
mkdir -p gpurun_out
free -g | head -2 > gpurun_out/${TAG}_host.txt; nproc >> gpurun_out/${TAG}_host.txt
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | tail -80 > gpurun_out/${TAG}_pytest_gpu.txt
timeout 300 python bench.py --steps 200 --warmup 20 2> gpurun_out/${TAG}_bench.err | tail -1 > gpurun_out/${TAG}_bench.json
for c in 1 2; do timeout 300 python bench.py --config $c --steps 100 --warmup 10 2> gpurun_out/${TAG}_bench_c$c.err | tail -1 > gpurun_out/${TAG}_bench_c$c.json; done
for c in 4 5; do timeout 400 python bench.py --config $c --steps 50 --warmup 5 --no-cpu-baseline 2> gpurun_out/${TAG}_bench_c$c.err | tail -1 > gpurun_out/${TAG}_bench_c$c.json; done
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 2> gpurun_out/${TAG}_bench_ref.err | tail -1 > gpurun_out/${TAG}_bench_ref.json
timeout 900 python tools/soup_capacity.py 1 4 16 64 > gpurun_out/${TAG}_soup_capacity.txt 2>&1
tail -30 gpurun_out/${TAG}_pytest_gpu.txt; cat gpurun_out/${TAG}_soup_capacity.txt | cut -c1-400; cat gpurun_out/${TAG}_host.txt
for f in bench bench_c1 bench_c2 bench_c4 bench_c5 bench_ref; do python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/${TAG}_$f.json").read())
    print("$f", j["value"], "e2e", j["e2e"]["value"], "launches", j.get("gpu_launches"), (j.get("roofline") or {}).get("frac"), (j.get("roofline_hbm") or {}).get("frac"))
except Exception as e:
    print("$f", "parse failed", e); print(open("gpurun_out/${TAG}_$f.err").read()[-1500:])
PY
done
