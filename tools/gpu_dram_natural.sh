#!/bin/bash
# DRAM bytes per launch with the L2 in its natural inter-kernel state: application replay, no cache control
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
COSK_NCU=1 timeout 1200 ncu --profile-from-start off --replay-mode application --cache-control none --clock-control none \
   --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file gpurun_out/dram_natural.csv \
   python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/dram_natural.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/dram_natural.log | cut -c1-200
