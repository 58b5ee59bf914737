#!/bin/bash
# usage: gpu_ncu.sh <kernel-regex> <tag>   -> gpurun_out/<tag>.ncu-rep (full set, 1 launch after 2 warm-up frames)
mkdir -p gpurun_out
python tools/profile_frame.py 3 --kernels 2>&1 | tail -40 | tee gpurun_out/kernels_$2.txt
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$1 --launch-skip ${SKIP:-2} -c ${COUNT:-1} -f -o gpurun_out/$2 python tools/profile_frame.py 3 > gpurun_out/ncu_$2.log 2>&1
tail -3 gpurun_out/ncu_$2.log
ls -la gpurun_out/*.ncu-rep
