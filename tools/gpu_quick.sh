#!/bin/bash
# quick round: GPU tests + cone-trace variant A/B (+ optional bench): gpu_quick.sh [bench]
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 | tee gpurun_out/pytest_gpu.txt
timeout 600 python tools/trace_variants.py 20 ${VARIANTS:-1 3 5 0} 2>&1 | tail -12 | tee gpurun_out/trace_variants.txt
python tools/profile_frame.py 3 --kernels 2>&1 | tail -40 | tee gpurun_out/kernels_frame.txt
if [ -n "$1" ]; then
timeout 600 python bench.py --steps ${STEPS:-200} --warmup 20 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -5 gpurun_out/bench.err; cut -c1-600 gpurun_out/bench.json
fi
