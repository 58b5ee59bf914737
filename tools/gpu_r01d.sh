#!/bin/bash
# round r01d: GPU tests, cone-trace variant A/B, bench, ncu launch list, ncu --set full of one whole frame
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 | tee gpurun_out/pytest_gpu.txt
timeout 600 python tools/trace_variants.py 20 2>&1 | tail -12 | tee gpurun_out/trace_variants.txt
python tools/profile_frame.py 3 --kernels 2>&1 | tail -40 | tee gpurun_out/kernels_frame.txt
timeout 600 python bench.py --steps ${STEPS:-200} --warmup 20 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -5 gpurun_out/bench.err
python - <<'PY'
import json
try:
    j=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
    print("value",j["value"],"e2e",j["e2e"]["value"],"launches",j["gpu_launches"],"clocks",j["clocks"])
    print("kernels_ms",j["kernels_ms"])
    print("voxel_passes",j["voxel_passes"])
    for r in j["roofline_passes"]: print(r["kernel"],r["ms"],r["achieved"],r["frac"])
except Exception as e: print("bench parse failed",e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --profile-frames 1 > gpurun_out/launches_bench.log 2>&1
tail -2 gpurun_out/launches_bench.log
# one whole frame (third of three), every kernel, full set
timeout 1500 ncu --set full --clock-control none --import-source on --launch-skip ${SKIP:-48} -c ${COUNT:-26} -f -o gpurun_out/frame_full python tools/profile_frame.py 3 > gpurun_out/ncu_frame_full.log 2>&1
tail -2 gpurun_out/ncu_frame_full.log
ncu -i gpurun_out/frame_full.ncu-rep --page raw --csv > gpurun_out/frame_full_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/frame_full_raw.csv > gpurun_out/frame_full_summary.txt 2>&1
ls -la gpurun_out/
