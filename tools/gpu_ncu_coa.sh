#!/bin/bash
# ncu --set full capture of the CoA-GCN kernels: tools/gpu_ncu_coa.sh <kernel regex> <count> <out name>
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
COSK_NCU=1 timeout 1500 ncu --profile-from-start off --set full --import-source on --clock-control none \
   -k regex:$1 -c $2 -o gpurun_out/$3 python bench.py --workload coa_gcn --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/$3.log 2>&1
echo "ncu rc=$?"
ls -la gpurun_out/ | grep $3
