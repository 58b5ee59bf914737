#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench1.json 2> gpurun_out/bench1.err
tail -5 gpurun_out/bench1.err
cat gpurun_out/bench1.json | head -c 6000
# launch list of a short run (shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 3 --warmup 3 --profile-frames 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/ncu_bench.log | head -c 1000
wc -l gpurun_out/launches_r1.csv
