#!/bin/bash
# One gpurun call: GPU parity tests (SIMT checker first, then tcgen05), smoke, bench, ncu launch list.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
run() { name=$1; shift; timeout $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "$name rc=$?" >> gpurun_out/summary.txt; }
run pytest_simt 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "simt"
run pytest_tc_blocks 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "auto and block"
run pytest_tc_models 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "not simt and not block"
run smoke 600 python __graft_entry__.py smoke
run bench_simt 900 python bench.py --kernel-path simt --steps 40 --warmup 3 --no-cpu-baseline
run bench_auto 900 python bench.py --steps 200 --warmup 8
run bench_auto_mod 900 python bench.py --workload cost_gcn_mod --steps 100 --warmup 8 --no-cpu-baseline
COSK_NCU=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file gpurun_out/launches_r1.csv python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
echo "ncu_list rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
for f in pytest_simt pytest_tc_blocks pytest_tc_models smoke bench_simt bench_auto bench_auto_mod; do echo "== $f"; tail -4 gpurun_out/$f.log | cut -c1-1500; done
