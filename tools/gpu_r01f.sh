#!/bin/bash
# round r01f: GPU tests, frame breakdown, bench, ncu --set full of one sparse frame
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.txt
python tools/profile_frame.py 4 --kernels 2>&1 | tail -32 | tee gpurun_out/kernels_frame.txt
timeout 600 python bench.py --steps ${STEPS:-200} --warmup 20 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -5 gpurun_out/bench.err
python - <<'PY'
import json
try:
    j=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
    print("value",j["value"],"e2e",j["e2e"]["value"],"launches",j["gpu_launches"],"clocks",j["clocks"])
    print("kernels_ms",j["kernels_ms"])
    print("voxel_passes",j["voxel_passes"])
except Exception as e: print("bench parse failed",e)
PY
timeout 1500 ncu --set full --clock-control none --import-source on --launch-skip ${SKIP:-66} -c ${COUNT:-24} -f -o gpurun_out/frame_full python tools/profile_frame.py 4 > gpurun_out/ncu_frame_full.log 2>&1
tail -2 gpurun_out/ncu_frame_full.log
ncu -i gpurun_out/frame_full.ncu-rep --page raw --csv > gpurun_out/frame_full_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/frame_full_raw.csv > gpurun_out/frame_full_summary.txt 2>&1
grep -E "^====|gpu__time_duration" gpurun_out/frame_full_summary.txt | paste - - | cut -c1-150
