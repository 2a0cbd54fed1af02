#!/bin/bash
# A/B: temporal conv on single CTAs (COSK_TCN_PAIR=0) vs CTA pairs (default): tests then benches
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "$name rc=$?" >> gpurun_out/summary.txt; }
COSK_TCN_PAIR=0 run pytest_gpu_single 1500 python -m pytest tests -m gpu -x -q
COSK_TCN_PAIR=0 run bench_single 900 python bench.py --steps 200 --warmup 8 --no-cpu-baseline
COSK_TCN_PAIR=7 run pytest_gpu_pair 1500 python -m pytest tests -m gpu -x -q
COSK_TCN_PAIR=7 run bench_pair 900 python bench.py --steps 200 --warmup 8 --no-cpu-baseline
COSK_TCN_PAIR=4 run bench_pair256 900 python bench.py --steps 200 --warmup 8 --no-cpu-baseline
COSK_TCN_PAIR=7 run bench_pair_mod 900 python bench.py --workload cost_gcn_mod --steps 100 --warmup 8 --no-cpu-baseline
cat gpurun_out/summary.txt
for f in pytest_gpu_single pytest_gpu_pair; do echo "== $f"; tail -15 gpurun_out/$f.log | cut -c1-300; done
