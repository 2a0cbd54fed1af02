"""One steady-state step of CoST-GCN NTU60 at 4096 streams in which every block fires (frame index = 0 mod 4), bracketed by
cudaProfilerStart / Stop: the capture target of `ncu --profile-from-start off --set full` (19 launches, every kernel of the
step once).  Source of profiles/ncu_full_main_summary.csv."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import continual_skeletons_b200 as cs  # noqa: E402

streams = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
torch.manual_seed(0)
m = cs.CoStGcn({"dataset_name": "dummy_ntu", "forward_mode": "frame", "time_chunk": 1})
frames = [torch.rand(streams, 3, 25, 2, device="cuda") for _ in range(4)]
for t in range(304):  # logits on frames 296, 300: steady state
    m.forward_step(frames[t % 4])
torch.cuda.synchronize()
l0 = m.launch_count()
torch.cuda.cudart().cudaProfilerStart()
out = m.forward_step(frames[0])  # frame 304: every block fires, logits are due
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
assert out is not None and m.device_error() == 0
print("launches in the profiled step:", m.launch_count() - l0, m.knobs()["blocks"])
