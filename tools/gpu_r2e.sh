#!/bin/bash
# Round 2: W_0 + residual folding in k_tc_gcn -- parity + A/B.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "$name rc=$?" >> gpurun_out/summary.txt; }
run pytest_all 1500 python -m pytest tests -m gpu -q -x
run bench_fold 600 python bench.py --steps 200 --warmup 8 --no-cpu-baseline
COSK_GCN_FOLD_UNIT=0 run bench_nofold 600 python bench.py --steps 200 --warmup 8 --no-cpu-baseline
run bench_fold_mod 600 python bench.py --workload cost_gcn_mod --steps 100 --warmup 8 --no-cpu-baseline
cat gpurun_out/summary.txt
grep -v "^E  " gpurun_out/pytest_all.log | tail -8 | cut -c1-300
python - <<'PY'
import json
for f in ('bench_fold','bench_nofold','bench_fold_mod'):
    txt=open(f'gpurun_out/{f}.log').read()
    for line in txt.split('\n'):
        if line.startswith('{'):
            d=json.loads(line)
            pb=d['kernel_time_per_block_ms']
            print(f, round(d['value']), 'e2e', round(d['e2e']['value']), 'ms/step', round(d['ms_per_step'],3), 'p50', round(d['p50_ms_per_step'],3), 'clk', d['clocks']['sm_mhz'], 'launches', d['gpu_launches'])
            print('   gcn', [round(b['gcn_ms']/max(b['gcn_n'],1),4) for b in pb]); print('   tcn', [round(b['tcn_ms']/max(b['tcn_n'],1),4) for b in pb]); print('   blk', [round(b['block_ms']/max(b['block_n'],1),4) for b in pb][1:4])
            print('   per-block hbm', [round(r['hbm_frac'],3) for r in d['per_block_roofline']], 'tensor(issued)', [round(r['tensor_frac_issued_3product'],3) for r in d['per_block_roofline']])
    if 'Traceback' in txt: print(txt[-800:])
PY
