#!/bin/bash
# Round 2: standalone pre-mix graph conv v2 (two-pass mix) -- parity with every width on it, then A/B benches + trace.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "$name rc=$?" >> gpurun_out/summary.txt; }
COSK_GCN_PREMIX=7 run pytest_p7 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "block_step_vs_golden or model_forward_steps or schedule_and_blocks or kinetics_skeleton or many_streams"
COSK_GCN_PREMIX=7 COSK_GCNP_STACKED=3 run pytest_p7s 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "block_step_vs_golden and auto"
COSK_GCN_PREMIX=7 COSK_FUSE_BLOCK=0 run pytest_p7nf 600 python -m pytest tests/test_gpu_scale.py -m gpu -q -x -s -k "distinct and cost_gcn"
run pytest_def 600 python -m pytest tests/test_gpu_scale.py -m gpu -q -x -s -k "distinct and cost_gcn"
for v in 0 2 6 7; do COSK_GCN_PREMIX=$v run bench_p$v 600 python bench.py --steps 100 --warmup 8 --no-cpu-baseline; done
COSK_GCN_PREMIX=6 COSK_GCNP_STACKED=3 run bench_p6s 600 python bench.py --steps 100 --warmup 8 --no-cpu-baseline
cat > /tmp/trace.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
import continual_skeletons_b200 as cs
for cin, cout in ((128, 128), (256, 256), (64, 128)):
    st = cs.CoStack([cs.BlockSpec(cin, cout, 1, True)], padding=4)
    N = 8192
    x = torch.rand(N, cin, 25, device='cuda')
    for t in range(6):
        st.forward_step(x)
    torch.cuda.synchronize()
    tr = st.trace_read(56)
    n = max(tr[37], 1)
    print(f"gcnp {cin}->{cout}: slots(cta0) {n} | mix per slot: wait_x {tr[32]//n} compute {tr[33]//n} wait_aslot {tr[34]//n} st+signal {tr[35]//n} total {tr[36]//n}"
          f" | mma per slot: wait_acc {tr[40]//n} wait_a {tr[41]//n} wait_w {tr[42]//n} issue {tr[43]//n} total {tr[44]//n}"
          f" | epi per slot: wait {tr[48]//n} work {tr[49]//n} total {tr[50]//n} | prod per slot: wait_x {tr[52]//n} wait_w {tr[53]//n} total {tr[54]//n}")
    print(st.knobs()["blocks"])
PY
COSK_TRACE=1 COSK_GCN_PREMIX=7 run trace_gcnp2 300 python /tmp/trace.py
cat gpurun_out/summary.txt
for f in pytest_p7 pytest_p7s pytest_p7nf pytest_def; do echo "== $f"; grep -v "^E  " gpurun_out/$f.log | tail -8 | cut -c1-300; done
cat gpurun_out/trace_gcnp2.log | cut -c1-700
python - <<'PY'
import json
for f in ('bench_p0','bench_p2','bench_p6','bench_p7','bench_p6s'):
    txt=open(f'gpurun_out/{f}.log').read()
    for line in txt.split('\n'):
        if line.startswith('{'):
            d=json.loads(line)
            pb=d['kernel_time_per_block_ms']
            print(f, round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'clk', d['clocks']['sm_mhz'], 'gcn', [round(b['gcn_ms']/max(b['gcn_n'],1),4) for b in pb], 'blk', [round(b['block_ms']/max(b['block_n'],1),4) for b in pb][1:4])
    if 'Traceback' in txt: print(txt[-800:])
PY
