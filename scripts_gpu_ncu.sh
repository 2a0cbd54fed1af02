#!/bin/bash
# ncu --set full capture of the tcgen05 kernels of one CoST-GCN step (phase with every block executing)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
COSK_NCU=1 timeout 1500 ncu --profile-from-start off --set full --import-source on --clock-control none \
   -k regex:k_tc_ -c 20 -o gpurun_out/prof_r1a python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
echo "ncu rc=$?"
ls -la gpurun_out/
tail -3 gpurun_out/ncu_full.log | cut -c1-400
