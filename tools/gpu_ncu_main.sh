#!/bin/bash
# ncu --set full (no source import, to keep the report small) of the tensor-core kernels of the default workload
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
COSK_NCU=1 timeout 1500 ncu --profile-from-start off --set full --clock-control none \
   -k regex:"k_tc_" -c ${1:-20} -o gpurun_out/main_full python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/main_full.log 2>&1
echo "ncu rc=$?"
ls -la gpurun_out/ | grep main_full
