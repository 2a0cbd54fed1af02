#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for mode in 0 1; do
COSK_TCN_REVERSE=$mode COSK_NCU=1 timeout 900 ncu --profile-from-start off --cache-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none --csv \
   --log-file gpurun_out/dram_nocc_rev$mode.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_dram$mode.log 2>&1
echo "ncu rev=$mode rc=$?"
done
