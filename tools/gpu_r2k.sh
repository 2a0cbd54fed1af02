#!/bin/bash
# Round 2: k_tc_gcnt with register reallocation (setmaxnreg), packed stores A/B, phase trace.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "$name rc=$?" >> gpurun_out/summary.txt; }
run pytest_gcnt 240 python -m pytest tests/test_gpu_parity.py -x -q -m gpu --timeout 40 -k "block_step_vs_golden or forward_steps_vs_reference or kinetics or distinct or ragged"
if ! grep -q "pytest_gcnt rc=0" gpurun_out/summary.txt; then tail -30 gpurun_out/pytest_gcnt.log | cut -c1-400; exit 1; fi
COSK_GCNT_PACK=0 run pytest_gcnt_nopack 240 python -m pytest tests/test_gpu_parity.py -x -q -m gpu --timeout 40 -k "block_step_vs_golden or kinetics"
run bench_pack 200 python bench.py --steps 100 --warmup 8 --no-cpu-baseline
COSK_GCNT_PACK=0 run bench_nopack 200 python bench.py --steps 100 --warmup 8 --no-cpu-baseline
COSK_GCN_T=0 run bench_gcn_old 200 python bench.py --steps 100 --warmup 8 --no-cpu-baseline
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
    print(f"k_tc_gcnt {cin}->{cout} pack={os.environ.get('COSK_GCNT_PACK','1')} ({st.knobs()['blocks'][0]['gcn']}): per chunk-item (cycles): epilogue wait_acc {tr[0]//n} load+mix {tr[1]//n} store {tr[2]//n} total {tr[3]//n} | "
          f"mma wait_acc_free {tr[8]//n} wait_x {tr[9]//n} wait_w {tr[10]//n} total {tr[11]//n} | producer wait_x_free {tr[12]//n} wait_w_free {tr[13]//n} | chunk-items {n}")
PY
COSK_TRACE=1 run trace_gcnt 300 python /tmp/trace.py
COSK_TRACE=1 COSK_GCNT_PACK=0 run trace_gcnt_nopack 300 python /tmp/trace.py
cat gpurun_out/summary.txt
tail -3 gpurun_out/pytest_gcnt.log gpurun_out/pytest_gcnt_nopack.log
cat gpurun_out/trace_gcnt.log gpurun_out/trace_gcnt_nopack.log | cut -c1-500
python - <<'PY'
import json
for f in ('bench_pack','bench_nopack','bench_gcn_old'):
    txt=open(f'gpurun_out/{f}.log').read()
    for line in txt.split('\n'):
        if line.startswith('{'):
            d=json.loads(line)
            pb=d['kernel_time_per_block_ms']
            print(f, round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'p50', round(d.get('p50_ms_per_step'),3), d.get('clocks',{}).get('sm_mhz'), 'gcn/launch', [round(b['gcn_ms']/max(b['gcn_n'],1),4) for b in pb])
    if 'Traceback' in txt: print(txt[-1500:])
PY
