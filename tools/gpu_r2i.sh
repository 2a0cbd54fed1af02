#!/bin/bash
# Round 2: artifacts -- benches (four workloads + reference arm), benchmark scripts, launch list, ncu --set full of one step, traces.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "$name rc=$?" >> gpurun_out/summary.txt; }
run bench_auto 900 python bench.py --steps 200 --warmup 8
run bench_auto_mod 900 python bench.py --workload cost_gcn_mod --steps 100 --warmup 8 --no-cpu-baseline
run bench_coa 900 python bench.py --workload coa_gcn --steps 100 --warmup 8 --no-cpu-baseline
run bench_cos 900 python bench.py --workload cos_tr --streams 2048 --steps 100 --warmup 8 --no-cpu-baseline
run bench_256 900 python bench.py --streams 256 --steps 200 --warmup 8 --no-cpu-baseline
run bench_reference 900 python bench.py --impl reference --steps 3 --warmup 1
run bench_script_ntu 900 python scripts/benchmark_all_ntu60.py
run bench_script_kin 900 python scripts/benchmark_all_kinetics.py
COSK_NCU=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file gpurun_out/launches.csv python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
echo "ncu_list rc=$?" >> gpurun_out/summary.txt
timeout 1200 ncu --profile-from-start off --set full --clock-control none -c 19 -o gpurun_out/main_full python tools/ncu_one_step.py > gpurun_out/main_full.log 2>&1
echo "ncu_full rc=$?" >> gpurun_out/summary.txt
cat > /tmp/trace.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
import continual_skeletons_b200 as cs
for cin, cout in ((64, 128), (128, 128), (256, 256)):
    st = cs.CoStack([cs.BlockSpec(cin, cout, 1, True)], padding=4)
    N = 8192
    x = torch.rand(N, cin, 25, device='cuda')
    for t in range(6):
        st.forward_step(x)
    torch.cuda.synchronize()
    tr = st.trace_read(24)
    n = max(tr[6], 1)
    print(f"k_tc_gcn {cin}->{cout} ({st.knobs()['blocks'][0]['gcn']}): per item (cycles): drain wait_acc {tr[0]//n} tmem+fold {tr[1]//n} wait_buf {tr[2]//n} write {tr[3]//n} total {tr[5]//n} | "
          f"mix wait_planes {tr[8]//n} gather+store {tr[9]//n} | mma wait_acc_free {tr[16]//n} wait_operands {tr[17]//n} total {tr[18]//n} | items {n}")
PY
COSK_TRACE=1 run trace_gcn_old 300 python /tmp/trace.py
ls -la gpurun_out/main_full.ncu-rep >> gpurun_out/summary.txt
du -sm gpurun_out >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
cat gpurun_out/trace_gcn_old.log | cut -c1-400
cat gpurun_out/bench_script_ntu.log gpurun_out/bench_script_kin.log | cut -c1-260
python - <<'PY'
import json
for f in ('bench_auto','bench_auto_mod','bench_coa','bench_cos','bench_256','bench_reference'):
    txt=open(f'gpurun_out/{f}.log').read()
    for line in txt.split('\n'):
        if line.startswith('{'):
            d=json.loads(line)
            print(f, round(d['value']), 'e2e', round(d['e2e']['value']), 'ms/step', round(d['ms_per_step'],4), 'p50', d.get('p50_ms_per_step'), 'launches', d.get('gpu_launches'), d.get('clocks'))
            if 'per_block_roofline' in d: print('   per-block hbm', [round(r['hbm_frac'],3) for r in d['per_block_roofline']], 'step', round(d['step_roofline']['hbm_frac'],3), 'dominant', d['roofline']['kernel'], round(d['roofline']['frac'],3))
    if 'Traceback' in txt: print(txt[-800:])
PY
