#!/bin/bash
# tests + bench + per-kernel frame breakdown + optional ncu capture: gpu_round.sh [kernel-regex tag]
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 | tee gpurun_out/pytest_gpu.txt
python tools/profile_frame.py 3 --kernels 2>&1 | tail -40 | tee gpurun_out/kernels_frame.txt
timeout 600 python bench.py --steps ${STEPS:-100} --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -5 gpurun_out/bench.err
python - <<'PY'
import json
try:
    j=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
    print("value",j["value"],"e2e",j["e2e"]["value"],"launches",j["gpu_launches"],"clocks",j["clocks"])
    print("kernels_ms",j["kernels_ms"])
    print("passes_ms",j["passes_ms"])
    print("voxel_passes",j["voxel_passes"])
    for r in j["roofline_passes"]: print(r["kernel"],r["ms"],r["achieved"],r["frac"])
    print("cpu",j["cpu_baseline"])
except Exception as e: print("bench parse failed",e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --profile-frames 1 > gpurun_out/launches_bench.log 2>&1
tail -2 gpurun_out/launches_bench.log
if [ -n "$1" ]; then
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:$1 --launch-skip ${SKIP:-0} -c ${COUNT:-12} -f -o gpurun_out/$2 python tools/profile_frame.py 2 > gpurun_out/ncu_$2.log 2>&1
tail -2 gpurun_out/ncu_$2.log; ls -la gpurun_out/*.ncu-rep
fi
