#!/bin/bash
# compute-sanitizer memcheck over a small model run (all kernel variants incl. CTA pairs)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
import continual_skeletons_b200 as cs
torch.manual_seed(0)
for cls in (cs.CoStGcn, cs.CoStGcnMod, cs.CoAGcn, cs.CoSTr):
    m = cls({"dataset_name": "dummy_ntu"})
    x = torch.rand(7, 3, 30, 25, 2, device='cuda')   # 14 skeletons -> 3 tiles (odd: phantom tile in the pair kernels)
    for t in range(30):
        m.forward_step(x[:, :, t].contiguous())
    torch.cuda.synchronize()
    print(cls.__name__, 'device_error', hex(m.device_error()), m.tensor_core_blocks())
st = cs.CoStack([cs.BlockSpec(64, 128, 2, True), cs.BlockSpec(128, 256, 1, True)], padding=0, skeleton="kinetics")
x = torch.rand(9, 64, 25, 18, device='cuda')
y = st.forward_steps(x)
torch.cuda.synchronize()
print('stack', None if y is None else tuple(y.shape), hex(st.device_error()))
st = cs.CoStack([cs.BlockSpec(64, 128, 2, True), cs.BlockSpec(128, 128, 1, True)], padding=4, skeleton="kinetics", adaptive=True)
x = torch.rand(9, 64, 25, 18, device='cuda')
y = st.forward_steps(x)
torch.cuda.synchronize()
print('adaptive stack', None if y is None else tuple(y.shape), hex(st.device_error()), st.tensor_core_blocks())
PY
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python /tmp/san.py > gpurun_out/sanitizer.log 2>&1
echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|device_error|stack|Invalid|Out-of-range" gpurun_out/sanitizer.log | head -20
