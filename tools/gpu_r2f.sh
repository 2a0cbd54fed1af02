#!/bin/bash
# Round 2: pad_end flush tests + full suite.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "$name rc=$?" >> gpurun_out/summary.txt; }
run pytest_padend 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "pad_end"
run pytest_all 1500 python -m pytest tests -m gpu -q
run smoke 600 python __graft_entry__.py smoke
cat gpurun_out/summary.txt
for f in pytest_padend pytest_all smoke; do echo "== $f"; grep -v "^E  " gpurun_out/$f.log | tail -12 | cut -c1-400; done
