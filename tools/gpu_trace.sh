#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "$name rc=$?" >> gpurun_out/summary.txt; }
run pytest_gpu 1500 python -m pytest tests -m gpu -x -q
cat > /tmp/trace.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
import continual_skeletons_b200 as cs
for cin, cout in ((64, 64), (128, 128), (256, 256), (64, 128)):
    st = cs.CoStack([cs.BlockSpec(cin, cout, 1, True)], padding=4)
    N = 8192
    x = torch.rand(N, cin, 25, device='cuda')
    for t in range(6):
        st.forward_step(x)
    torch.cuda.synchronize()
    tr = st.trace_read(24)
    n = max(tr[6], 1)
    print(f"gcn {cin}->{cout}: per item (cycles): drain wait_acc {tr[0]//n} tmem+fold {tr[1]//n} wait_buf {tr[2]//n} write {tr[3]//n} total {tr[5]//n} | "
          f"mix wait_planes {tr[8]//n} gather+store {tr[9]//n} | mma wait_acc_free {tr[16]//n} wait_operands {tr[17]//n} total {tr[18]//n} | items {n}")
PY
COSK_TRACE=1 run trace 300 python /tmp/trace.py
run bench_auto 900 python bench.py --steps 200 --warmup 8 --no-cpu-baseline
run bench_auto_mod 900 python bench.py --workload cost_gcn_mod --steps 100 --warmup 8 --no-cpu-baseline
cat gpurun_out/summary.txt
tail -8 gpurun_out/pytest_gpu.log | cut -c1-300
cat gpurun_out/trace.log | cut -c1-400
