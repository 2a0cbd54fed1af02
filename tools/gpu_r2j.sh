#!/bin/bash
# Round 2: k_tc_gcnt (channel-major graph conv, adjacency mix in registers) -- parity subset, then A/B against k_tc_gcn.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "$name rc=$?" >> gpurun_out/summary.txt; }
run pytest_gcnt 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "block_step_vs_golden or forward_steps_vs_reference or kinetics or distinct or ragged"
run bench_gcnt 600 python bench.py --steps 100 --warmup 8 --no-cpu-baseline
COSK_GCN_T=0 run bench_gcn_old 600 python bench.py --steps 100 --warmup 8 --no-cpu-baseline
run bench_gcnt2 600 python bench.py --steps 100 --warmup 8 --no-cpu-baseline
cat gpurun_out/summary.txt
tail -5 gpurun_out/pytest_gcnt.log
python - <<'PY'
import json
for f in ('bench_gcnt','bench_gcn_old','bench_gcnt2'):
    txt=open(f'gpurun_out/{f}.log').read()
    for line in txt.split('\n'):
        if line.startswith('{'):
            d=json.loads(line)
            print(f, round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'p50', round(d.get('p50_ms_per_step'),3), d.get('clocks',{}).get('sm_mhz'))
            print('   per-block ms', [round(r['ms_per_block_step'],3) for r in d['per_block_roofline']])
            if 'kernel_ms' in d: print(d['kernel_ms'])
    if 'Traceback' in txt: print(txt[-1500:])
PY
