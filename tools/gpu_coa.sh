#!/bin/bash
# CoA-GCN bring-up: adaptive parity tests, then a short bench of the coa_gcn workload
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "$name rc=$?" >> gpurun_out/summary.txt; }
run pytest_coa 1500 python -m pytest tests -m gpu -x -q -k "${COA_K:-adaptive or coa}"
run bench_coa 900 python bench.py --workload ${COA_W:-coa_gcn} --steps ${COA_STEPS:-40} --warmup 4 --no-cpu-baseline ${COA_ARGS:-}
cat gpurun_out/summary.txt; tail -5 gpurun_out/pytest_coa.log | cut -c1-300; tail -2 gpurun_out/bench_coa.log | cut -c1-1500
