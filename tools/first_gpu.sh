#!/bin/bash
# first GPU call of the build: probe the box, then parity tests + smoke
mkdir -p gpurun_out
{ nvidia-smi; ls /root/reference 2>&1 | head -3; ls baseline/_ref 2>&1 | head -3; ldconfig -p | grep -iE 'libEGL|libGL|OSMesa'; nproc; free -g | head -2; du -sh assets/_baked; } > gpurun_out/probe.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -60 | tee gpurun_out/pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -20 | tee gpurun_out/smoke.txt
