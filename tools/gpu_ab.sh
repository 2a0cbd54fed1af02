#!/bin/bash
# A/B helper: tests, then bench with an environment knob off/on (usage: gpu_ab.sh KNOB)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "$name rc=$?" >> gpurun_out/summary.txt; }
run pytest_gpu 1500 python -m pytest tests -m gpu -x -q
for rep in 1 2; do
  env $1=0 timeout 900 python bench.py --steps 200 --warmup 8 --no-cpu-baseline > gpurun_out/bench_off$rep.log 2>&1; echo "off$rep rc=$?" >> gpurun_out/summary.txt
  env $1=1 timeout 900 python bench.py --steps 200 --warmup 8 --no-cpu-baseline > gpurun_out/bench_on$rep.log 2>&1; echo "on$rep rc=$?" >> gpurun_out/summary.txt
done
cat gpurun_out/summary.txt; tail -3 gpurun_out/pytest_gpu.log | cut -c1-200
