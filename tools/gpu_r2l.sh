#!/bin/bash
# Round 2: k_tc_gcnt validation (guarded), then the full suite, parity report, benches, launch list and ncu --set full export.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "$name rc=$?" >> gpurun_out/summary.txt; }
run pytest_gcnt 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu --timeout 60 -k "channel_major or block_step_vs_golden or kinetics"
if ! grep -q "pytest_gcnt rc=0" gpurun_out/summary.txt; then tail -40 gpurun_out/pytest_gcnt.log | cut -c1-400; exit 1; fi
run bench_auto 300 python bench.py --steps 200 --warmup 8
cat > /tmp/trace.py <<'PY'
import sys, os, torch
sys.path.insert(0, '.')
import continual_skeletons_b200 as cs
for cin, cout in ((64, 128), (128, 128), (128, 256)):
    st = cs.CoStack([cs.BlockSpec(cin, cout, 1, True)], padding=4)
    N = 4096
    x = torch.rand(N, cin, 25, device='cuda')
    for t in range(6):
        st.forward_step(x)
    torch.cuda.synchronize()
    tr = st.trace_read(16)
    n = max(tr[4], 1)
    print(f"k_tc_gcnt {cin}->{cout} ({st.knobs()['blocks'][0]['gcn']}): per chunk-item (cycles): epilogue wait_acc {tr[0]//n} load+mix {tr[1]//n} store {tr[2]//n} total {tr[3]//n} | "
          f"mma wait_acc_free {tr[8]//n} wait_x {tr[9]//n} wait_w {tr[10]//n} total {tr[11]//n} | producer wait_x_free {tr[12]//n} wait_w_free {tr[13]//n} | chunk-items {n}")
PY
COSK_TRACE=1 run trace_gcnt 120 python /tmp/trace.py
run pytest_all 900 python -m pytest tests -m gpu -q --timeout 180
run parity_report 600 python tools/gpu_parity_report.py
run bench_auto_mod 300 python bench.py --workload cost_gcn_mod --steps 100 --warmup 8 --no-cpu-baseline
run bench_coa 300 python bench.py --workload coa_gcn --steps 100 --warmup 8 --no-cpu-baseline
run bench_cos 300 python bench.py --workload cos_tr --streams 2048 --steps 100 --warmup 8 --no-cpu-baseline
run bench_256 300 python bench.py --streams 256 --steps 200 --warmup 8 --no-cpu-baseline
COSK_NCU=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file gpurun_out/launches.csv python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
echo "ncu_list rc=$?" >> gpurun_out/summary.txt
timeout 900 ncu --profile-from-start off --set full --clock-control none -c 19 -o /tmp/main_full python tools/ncu_one_step.py > gpurun_out/main_full.log 2>&1
echo "ncu_full rc=$?" >> gpurun_out/summary.txt
ncu -i /tmp/main_full.ncu-rep --page raw --csv > gpurun_out/main_full_raw.csv 2>/dev/null
cat gpurun_out/summary.txt
tail -n 4 gpurun_out/pytest_all.log | cut -c1-300
cat gpurun_out/trace_gcnt.log | cut -c1-420
python - <<'PY'
import json
for f in ('bench_auto','bench_auto_mod','bench_coa','bench_cos','bench_256'):
    txt=open(f'gpurun_out/{f}.log').read()
    for line in txt.split('\n'):
        if line.startswith('{'):
            d=json.loads(line)
            print(f, round(d['value']), 'e2e', round(d['e2e']['value']), 'ms/step', round(d['ms_per_step'],4), 'p50', round(d.get('p50_ms_per_step'),3), 'launches', d.get('gpu_launches'), d.get('clocks',{}).get('sm_mhz'))
            if 'per_block_roofline' in d: print('   per-block hbm', [round(r['hbm_frac'],3) for r in d['per_block_roofline']], 'step', round(d['step_roofline']['hbm_frac'],3), 'dominant', d['roofline']['kernel'], round(d['roofline']['frac'],3), d['roofline'].get('traffic'))
    if 'Traceback' in txt: print(txt[-800:])
PY
