#!/bin/bash
# Round 2: phase trace + ncu of the pre-mix graph conv, new scale tests.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout $1 "${@:2}" > gpurun_out/$name.log 2>&1; echo "$name rc=$?" >> gpurun_out/summary.txt; }
cat > /tmp/trace.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
import continual_skeletons_b200 as cs
for cin, cout in ((64, 64), (128, 128), (256, 256)):
    st = cs.CoStack([cs.BlockSpec(cin, cout, 1, True)], padding=4)
    N = 8192
    x = torch.rand(N, cin, 25, device='cuda')
    for t in range(6):
        st.forward_step(x)
    torch.cuda.synchronize()
    tr = st.trace_read(56)
    n = max(tr[37], 1)
    tiles = (N + 4) // 5 // 148 + 1
    print(f"gcnp {cin}->{cout}: slots(cta0) {n} | mix per slot: wait_x {tr[32]//n} compute {tr[33]//n} wait_aslot {tr[34]//n} st+signal {tr[35]//n} total {tr[36]//n}"
          f" | mma per slot: wait_acc {tr[40]//n} wait_a {tr[41]//n} wait_w {tr[42]//n} issue {tr[43]//n} total {tr[44]//n}"
          f" | epi per slot: wait {tr[48]//n} work {tr[49]//n} total {tr[50]//n} | prod per slot: wait_x {tr[52]//n} wait_w {tr[53]//n} total {tr[54]//n}")
PY
COSK_TRACE=1 run trace_gcnp 300 python /tmp/trace.py
run pytest_scale 1500 python -m pytest tests/test_gpu_scale.py -m gpu -x -q -s
COSK_NCU=1 timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"k_tc_gcnp" -c 12 -o gpurun_out/gcnp_full python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/gcnp_full.log 2>&1
echo "ncu rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
cat gpurun_out/trace_gcnp.log | cut -c1-700
tail -40 gpurun_out/pytest_scale.log | cut -c1-300
