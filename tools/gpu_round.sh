#!/bin/bash
# Round capture: GPU tests, bench (+ reference arm), per-kernel frame breakdown, ncu launch list of the bench command,
# ncu --set full of one (sparse) frame -> summary + traffic.json.   usage: gpu_round.sh [tag]
TAG=${1:-round}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 | tee gpurun_out/${TAG}_pytest_gpu.txt
python tools/profile_frame.py 4 --kernels 2>&1 | tail -32 | tee gpurun_out/${TAG}_kernels_frame.txt
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -5 gpurun_out/${TAG}_bench.err
python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
    print("value",j["value"],"e2e",j["e2e"]["value"],"launches",j["gpu_launches"],"clocks",j["clocks"])
    print("kernels_ms",j["kernels_ms"])
    print("voxel_passes",j["voxel_passes"]); print("roofline",j["roofline"]); print("cpu",j["cpu_baseline"])
except Exception as e: print("bench parse failed",e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --profile-frames 1 > gpurun_out/${TAG}_launches_bench.log 2>&1
tail -2 gpurun_out/${TAG}_launches_bench.log
timeout 1500 ncu --set full --clock-control none --import-source on --launch-skip ${SKIP:-60} -c ${COUNT:-22} -f -o gpurun_out/${TAG}_frame_full python tools/profile_frame.py 4 > gpurun_out/${TAG}_ncu_frame_full.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_frame_full.log
ncu -i gpurun_out/${TAG}_frame_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_frame_full_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/${TAG}_frame_full_raw.csv > gpurun_out/${TAG}_frame_full_summary.txt 2>&1
python tools/ncu_traffic.py gpurun_out/${TAG}_frame_full_raw.csv gpurun_out/${TAG}_traffic.json
grep -E "^====|gpu__time_duration" gpurun_out/${TAG}_frame_full_summary.txt | paste - - | cut -c1-150
