#!/bin/bash
# round 2, first GPU call: every GPU test (no -x), voxel-view diagnostic, split cone-trace A/B, bench lines
TAG=r02a
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -s 2>&1 | tail -60 > gpurun_out/${TAG}_pytest_gpu.txt
timeout 200 python tools/diag_voxel_view.py > gpurun_out/${TAG}_diag_voxel_view.txt 2>&1
timeout 200 python tools/diag_voxel_view.py 333 251 >> gpurun_out/${TAG}_diag_voxel_view.txt 2>&1
timeout 500 python tools/trace_variants.py 20 0 64 192 320 448 576 704 2>&1 | tail -12 > gpurun_out/${TAG}_trace_split_ab.txt
timeout 300 python bench.py --steps 200 --warmup 20 2> gpurun_out/${TAG}_bench.err | tail -1 > gpurun_out/${TAG}_bench.json
for v in 192 448; do VCT_TRACE_VARIANT=$v timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline 2> gpurun_out/${TAG}_bench_v$v.err | tail -1 > gpurun_out/${TAG}_bench_v$v.json; done
tail -5 gpurun_out/${TAG}_pytest_gpu.txt; cat gpurun_out/${TAG}_diag_voxel_view.txt | head -30; cat gpurun_out/${TAG}_trace_split_ab.txt
