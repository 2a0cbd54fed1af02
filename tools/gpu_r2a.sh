#!/bin/bash
# Round 2, first call: pre-mix graph conv bring-up -- parity subset, then A/B benches.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "$name rc=$?" >> gpurun_out/summary.txt; }
run pytest_premix 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "block_step_vs_golden or model_forward_steps or schedule_and_blocks or kinetics_skeleton or many_streams"
COSK_GCNP_STACKED=0 run pytest_premix_ns 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "block_step_vs_golden and auto"
COSK_GCNP_IDENTITY_MMA=0 run pytest_premix_epi 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "block_step_vs_golden and auto"
COSK_GCN_PREMIX=0 run bench_old 600 python bench.py --steps 100 --warmup 8 --no-cpu-baseline
run bench_new 600 python bench.py --steps 100 --warmup 8 --no-cpu-baseline
COSK_GCNP_STACKED=3 run bench_new_st3 600 python bench.py --steps 100 --warmup 8 --no-cpu-baseline
COSK_GCNP_STACKED=0 run bench_new_st0 600 python bench.py --steps 100 --warmup 8 --no-cpu-baseline
COSK_GCNP_IDENTITY_MMA=0 run bench_new_id0 600 python bench.py --steps 100 --warmup 8 --no-cpu-baseline
COSK_GCNP_IDENTITY_MMA=7 run bench_new_id7 600 python bench.py --steps 100 --warmup 8 --no-cpu-baseline
COSK_GCN_PREMIX=3 run bench_new_p3 600 python bench.py --steps 100 --warmup 8 --no-cpu-baseline
cat gpurun_out/summary.txt
for f in pytest_premix pytest_premix_ns pytest_premix_epi; do echo "== $f"; tail -15 gpurun_out/$f.log | cut -c1-300; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_*.log')):
    if not any(k in f for k in ('bench_old','bench_new')): continue
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line)
            pb=d['kernel_time_per_block_ms']
            print(f, round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'gcn ms/launch', [round(b['gcn_ms']/max(b['gcn_n'],1),4) for b in pb], 'tcn', [round(b['tcn_ms']/max(b['tcn_n'],1),4) for b in pb])
PY
