#!/bin/bash
# Round-end evidence on one B200: the whole GPU test-suite as the driver runs it, smoke(), the bench line of every configuration, the reference arm,
# and the ncu captures behind profiles/traffic.json and the launch list.   usage: gpurun --timeout 2400 -- 'bash tools/gpu_final.sh [tag]'
TAG=${1:-r02z}
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider ) > gpurun_out/${TAG}_pytest_gpu.txt 2>&1
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) > gpurun_out/${TAG}_smoke.txt 2>&1
timeout 600 python bench.py 2> gpurun_out/${TAG}_bench.err | tail -1 > gpurun_out/${TAG}_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2> gpurun_out/${TAG}_bench_reference.err | tail -1 > gpurun_out/${TAG}_bench_reference.json
for C in 1 2 5 4; do
timeout 400 python bench.py --config $C --steps $([ $C -le 2 ] && echo 200 || echo 50) --warmup 10 --no-cpu-baseline 2> gpurun_out/${TAG}_bench_config$C.err | tail -1 > gpurun_out/${TAG}_bench_config$C.json
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/${TAG}_launches.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__t_output_wavefronts_pipe_tex_mem_texture.sum --clock-control none --csv --page raw --log-file gpurun_out/${TAG}_frame_metrics.csv python tools/profile_frame.py 3 > gpurun_out/${TAG}_frame_metrics.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"k_cone_trace|k_voxel_tiles|k_voxel_bin|k_voxel_resolve|k_inject_linear|k_mip_chain|k_frame_begin" --launch-skip 14 -c 7 --csv --page raw --log-file gpurun_out/${TAG}_frame_full_raw.csv python tools/profile_frame.py 3 > gpurun_out/${TAG}_frame_full.log 2>&1
python tools/ncu_summary.py gpurun_out/${TAG}_frame_full_raw.csv > gpurun_out/${TAG}_frame_ncu_full_summary.txt 2>&1
tail -4 gpurun_out/${TAG}_pytest_gpu.txt; tail -3 gpurun_out/${TAG}_smoke.txt
for f in bench bench_reference bench_config1 bench_config2 bench_config5 bench_config4; do python -c "
import json; j=json.loads(open('gpurun_out/${TAG}_$f.json').read()); print('$f', j.get('value'), (j.get('e2e') or {}).get('value'), j.get('roofline',{}).get('frac'), (j.get('cpu_baseline') or {}).get('value'))" || tail -5 gpurun_out/${TAG}_$f.err; done
wc -l gpurun_out/${TAG}_launches.csv gpurun_out/${TAG}_frame_metrics.csv gpurun_out/${TAG}_frame_full_raw.csv; head -30 gpurun_out/${TAG}_frame_ncu_full_summary.txt
