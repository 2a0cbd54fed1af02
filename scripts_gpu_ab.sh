#!/bin/bash
# A/B of launch/cache features: tests with everything on, benches with PDL off/on
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "$name rc=$?" >> gpurun_out/summary.txt; }
run pytest_gpu 1500 python -m pytest tests -m gpu -x -q
run smoke 600 python __graft_entry__.py smoke
COSK_PDL=0 run bench_nopdl 900 python bench.py --steps 200 --warmup 8 --no-cpu-baseline
run bench_pdl 900 python bench.py --steps 200 --warmup 8 --no-cpu-baseline
run bench_pdl_mod 900 python bench.py --workload cost_gcn_mod --steps 100 --warmup 8 --no-cpu-baseline
COSK_NCU=1 timeout 900 ncu --profile-from-start off --set full --clock-control none \
   -k regex:"k_tc_tcn" -c 10 -o gpurun_out/prof_tcn python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_tcn.log 2>&1
echo "ncu_tcn rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
for f in pytest_gpu smoke; do echo "== $f"; tail -12 gpurun_out/$f.log | cut -c1-400; done
