#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
cat > /tmp/trace.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
import continual_skeletons_b200 as cs
names = ["wait_acc", "tmem_ld", "exchange", "gather", "store", "total", "items"]
for cin, cout in ((64, 64), (128, 128), (256, 256), (64, 128)):
    st = cs.CoStack([cs.BlockSpec(cin, cout, 1, True)], padding=4)
    N = 8192
    x = torch.rand(N, cin, 25, device='cuda')
    for t in range(6):
        st.forward_step(x)
    torch.cuda.synchronize()
    tr = st.trace_read(24)
    for s in (0, 1):
        d = dict(zip(names, tr[s * 8: s * 8 + 7]))
        n = max(d["items"] // 2, 1)
        print(f"gcn {cin}->{cout} set{s}: per item (cycles):", {k: round(v / n) for k, v in d.items() if k != "items"}, "items/set", n)
    print(f"   mma thread: wait_acc_free {tr[16]} wait_operands {tr[17]} total {tr[18]}")
PY
COSK_TRACE=1 timeout 300 python /tmp/trace.py > gpurun_out/trace.log 2>&1
echo "trace rc=$?"; cat gpurun_out/trace.log | cut -c1-400
COSK_TCN_PAIR=6 timeout 600 python bench.py --steps 200 --warmup 8 --no-cpu-baseline > gpurun_out/bench_hint.log 2>&1; echo "bench rc=$?"
