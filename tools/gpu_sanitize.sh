#!/bin/bash
# compute-sanitizer over small whole frames (the new build's race / memory checking): memcheck and racecheck in the
# deterministic (per-voxel lists + resolve), free-running CAS and warp/temporal modes; with >= 2 GPUs also the sharded frame
# (peer stores of exchange.cu) under memcheck.   usage: gpurun --timeout 1500 -- 'bash tools/gpu_sanitize.sh [tag]'
TAG=${1:-r02}
OUT=gpurun_out/${TAG}_sanitizer.txt
mkdir -p gpurun_out; : > $OUT
for mode in ${MODES:-det cas warp}; do
  for tool in ${TOOLS:-memcheck racecheck}; do
    echo "==== compute-sanitizer --tool $tool : sanitize_frame.py $mode" >> $OUT
    timeout 600 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 9 python tools/sanitize_frame.py $mode 2 2>&1 | grep -v "^=========  *$" | tail -25 >> $OUT
    echo "exit code ${PIPESTATUS[0]}" >> $OUT
  done
done
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  echo "==== compute-sanitizer --tool memcheck : vct_headless --gpus 2 (single process, peer stores)" >> $OUT
  python tools/pack_scene.py room /tmp/room.vcts >> $OUT 2>&1
  timeout 900 compute-sanitizer --tool memcheck --print-limit 20 --error-exitcode 9 vct_b200/lib/vct_headless /tmp/room.vcts --dim 64 --levels 5 --size 320x240 --shadow 512 --volume -1.5 1.5 --gpus 2 --frames 3 2>&1 | tail -25 >> $OUT
  echo "exit code ${PIPESTATUS[0]}" >> $OUT
fi
tail -60 $OUT
