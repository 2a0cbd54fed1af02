"""Print the measured logit parity of the CUDA path against the golden fixtures (reference blocks) and
the step oracle for the north-star clip (N=2, C=3, T=300, V=25, M=2).  Run on the GPU box."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import continual_skeletons_b200 as cs  # noqa: E402
from oracle import step, weights  # noqa: E402

x = weights.make_input((2, 3, 300, 25, 2), seed=11)
for cls, arch_fn, tag in ((cs.CoStGcn, weights.cost_gcn_arch, "cost_gcn"), (cs.CoStGcnMod, weights.cost_gcn_mod_arch, "cost_gcn_mod")):
    gold = np.load(os.path.join(ROOT, "tests", "golden", tag + ".npz"))
    for rnd in (False, True):
        for path in ("auto", "simt"):
            arch = arch_fn()
            sd = weights.make_state_dict(arch, seed=8 if rnd else 7, randomize=rnd)
            m = cls({"dataset_name": "dummy_ntu", "kernel_path": path})
            m.load_state_dict(m.map_state_dict(sd), strict=True)
            out = m.forward_steps(x.cuda()).cpu()
            want = torch.from_numpy(gold[f"{tag}_co_logits{'_rnd' if rnd else ''}"])
            ref = step.StepModel(sd, arch)
            with torch.no_grad():
                so = ref.forward_steps(x)
            print(f"{tag:13s} weights={'randomised-BN' if rnd else 'reference-init':15s} path={path:4s} max|logit|={float(want.abs().max()):7.2f} "
                  f"max|d| vs reference blocks={float((out - want).abs().max()):.2e}  vs step oracle={float((out - so).abs().max()):.2e}  "
                  f"argmax equal={bool(torch.equal(out.argmax(1), want.argmax(1)))}  oracle vs reference blocks={float((so - want).abs().max()):.2e}")
