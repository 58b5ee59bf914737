#!/bin/bash
# on the GPU box (tools/hw_smoke.py explains): one vct_headless run per configuration, ~8 s in all
E=vct_b200/lib/vct_headless; S=gpurun_in/room.vcts
A="--dim 64 --levels 6 --size 256x256 --shadow 1024 --frames 2 --eye 1.1 0.3 1.2 --front -0.65 -0.25 -0.72 --volume -1.5 1.5"
run() { n=$1; shift; timeout 8 $E $S $A "$@" --out gpurun_out/hw_$n.ppm > gpurun_out/hw_$n.json 2> gpurun_out/hw_$n.err; echo "$n rc=$?"; }
run plain
run tess_warp --tesselation --atomic-max --tesselation-warp
run raster_tess_warp --tesselation-warp
run view_tess_warp --tesselation --atomic-max --tesselation-warp --view voxels --miplevel 0
run both_warps --warp-texture --warp-voxels
