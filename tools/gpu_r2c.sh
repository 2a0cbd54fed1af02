#!/bin/bash
# Round 2: fused block kernel bring-up -- parity, then A/B bench and phase trace.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "$name rc=$?" >> gpurun_out/summary.txt; }
run pytest_fuse1 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "block_step_vs_golden and auto"
run pytest_fuse2 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "model_forward_steps or schedule_and_blocks or many_streams or state_lifecycle or single_stream"
run pytest_fuse3 900 python -m pytest tests/test_gpu_scale.py -m gpu -q -x -s -k "distinct and cost_gcn"
run bench_fuse 600 python bench.py --steps 100 --warmup 8 --no-cpu-baseline
COSK_FUSE_BLOCK=0 run bench_nofuse 600 python bench.py --steps 100 --warmup 8 --no-cpu-baseline
run bench_fuse_mod 600 python bench.py --workload cost_gcn_mod --steps 100 --warmup 8 --no-cpu-baseline
cat > /tmp/trace.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
import continual_skeletons_b200 as cs
st = cs.CoStack([cs.BlockSpec(64, 64, 1, True)], padding=4)
N = 8192
x = torch.rand(N, 64, 25, device='cuda')
for t in range(8):
    st.forward_step(x)
torch.cuda.synchronize()
tr = st.trace_read(56)
n = max(tr[37], 1)
print(f"block64: tiles(cta0) {n} | mix per tile: wait_x {tr[32]//n} compute {tr[33]//n} wait_slot {tr[34]//n} st+signal {tr[35]//n} total {tr[36]//n}"
      f" | mma per tile: wait_gacc {tr[40]//n} wait_mix {tr[41]//n} wait_w {tr[42]//n} wait_tacc {tr[43]//n} wait_hist {tr[44]//n} wait_tap {tr[45]//n} total {tr[46]//n}"
      f" | epi per tile: wait_g {tr[48]//n} g_work {tr[49]//n} wait_t {tr[50]//n} t_work {tr[51]//n} | prod per tile: wait_a {tr[52]//n} wait_b {tr[53]//n} total {tr[54]//n}")
print(st.knobs()["blocks"])
PY
COSK_TRACE=1 run trace_block 300 python /tmp/trace.py
cat gpurun_out/summary.txt
for f in pytest_fuse1 pytest_fuse2 pytest_fuse3; do echo "== $f"; tail -12 gpurun_out/$f.log | cut -c1-400; done
cat gpurun_out/trace_block.log | cut -c1-900
python - <<'PY'
import json
for f in ('bench_fuse','bench_nofuse','bench_fuse_mod'):
    for line in open(f'gpurun_out/{f}.log'):
        if line.startswith('{'):
            d=json.loads(line)
            pb=d['kernel_time_per_block_ms']
            print(f, round(d['value']), 'e2e', round(d['e2e']['value']), 'ms/step', round(d['ms_per_step'],3), 'p50', round(d['p50_ms_per_step'],3), 'launches', d['gpu_launches'], d['clocks'])
            print('   gcn', [round(b['gcn_ms']/max(b['gcn_n'],1),4) for b in pb]); print('   tcn', [round(b['tcn_ms']/max(b['tcn_n'],1),4) for b in pb]); print('   blk', [round(b['block_ms']/max(b['block_n'],1),4) for b in pb])
            print('   per-block hbm frac', [round(r['hbm_frac'],3) for r in d['per_block_roofline']])
    print(open(f'gpurun_out/{f}.log').read()[-600:] if 'Traceback' in open(f'gpurun_out/{f}.log').read() else '')
PY
