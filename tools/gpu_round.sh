#!/bin/bash
# One gpurun call: full GPU test suite, smoke, benches (both variants), ncu launch list, ncu full of the graph conv.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "$name rc=$?" >> gpurun_out/summary.txt; }
run pytest_gpu 1500 python -m pytest tests -m gpu -x -q
run smoke 600 python __graft_entry__.py smoke
run bench_auto 900 python bench.py --steps 200 --warmup 8
run bench_auto_mod 900 python bench.py --workload cost_gcn_mod --steps 100 --warmup 8 --no-cpu-baseline
run bench_coa 900 python bench.py --workload coa_gcn --steps 100 --warmup 8 --no-cpu-baseline
run bench_cos 900 python bench.py --workload cos_tr --streams 2048 --steps 100 --warmup 8 --no-cpu-baseline
run bench_script 600 python scripts/benchmark_all_ntu60.py
run bench_reference 900 python bench.py --impl reference --steps 3 --warmup 1
COSK_NCU=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file gpurun_out/launches.csv python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
echo "ncu_list rc=$?" >> gpurun_out/summary.txt
COSK_NCU=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file gpurun_out/launches_coa.csv python bench.py --workload coa_gcn --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list_coa.log 2>&1
echo "ncu_list_coa rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
for f in pytest_gpu smoke; do echo "== $f"; tail -12 gpurun_out/$f.log | cut -c1-400; done
