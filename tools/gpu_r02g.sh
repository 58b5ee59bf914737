#!/bin/bash
# ncu --set full of the voxel-pass kernels of a steady-state (sparse) frame + the launch list of a bench run
TAG=r02g
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_voxel|k_inject|k_mip|k_frame_begin|k_clear" --launch-skip 16 -c 24 -f -o gpurun_out/${TAG}_voxel python tools/profile_frame.py 4 > gpurun_out/${TAG}_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_ncu.log
ncu -i gpurun_out/${TAG}_voxel.ncu-rep --page raw --csv > gpurun_out/${TAG}_voxel_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/${TAG}_voxel_raw.csv "" 9 > gpurun_out/${TAG}_voxel_summary.txt 2>&1
grep -E "^====|gpu__time_duration|registers_per_thread|warps_active|dram__bytes_read|dram__bytes_write|inst_executed.sum|stalls" gpurun_out/${TAG}_voxel_summary.txt | tail -80
ls -la gpurun_out/${TAG}_voxel.ncu-rep
