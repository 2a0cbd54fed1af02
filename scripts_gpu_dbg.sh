#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
cat > /tmp/repro.py <<'PY'
import torch, sys
sys.path.insert(0, '.')
import continual_skeletons_b200 as cs
st = cs.CoStack([cs.BlockSpec(64, 64, 1, True)], padding=4)
x = torch.rand(6, 64, 12, 25, device='cuda')   # 6 skeletons -> 2 tiles
for t in range(12):
    o = st.forward_step(x[:, :, t].contiguous())
torch.cuda.synchronize()
print('device_error', hex(st.device_error()), 'ok', o.shape, float(o.abs().sum()))
PY
COSK_TCN_PAIR=7 timeout 300 compute-sanitizer --tool memcheck --print-limit 10 python /tmp/repro.py > gpurun_out/sanitizer.log 2>&1
echo "sanitizer rc=$?"
tail -60 gpurun_out/sanitizer.log | cut -c1-300
