#!/bin/bash
# Round 2: time-batched forward_steps -- new tests, then the full suite, then a bench sanity run.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "$name rc=$?" >> gpurun_out/summary.txt; }
run pytest_batched 900 python -m pytest tests/test_gpu_scale.py -m gpu -q -x -k "time_batched"
run pytest_all 1800 python -m pytest tests -m gpu -q
run bench 600 python bench.py --steps 100 --warmup 8 --no-cpu-baseline
cat gpurun_out/summary.txt
for f in pytest_batched pytest_all; do echo "== $f"; grep -v "^E   " gpurun_out/$f.log | tail -25 | cut -c1-300; done
grep '^{' gpurun_out/bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print(round(d['value']), d['ms_per_step'], d['gpu_launches'], d['clocks'])"
