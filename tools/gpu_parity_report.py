"""Measured logit parity of the CUDA path, UNSCALED, for all four models: against the golden fixtures made from the
reference's own blocks and against the step oracle, for the north-star clip (N=2, C=3, T=300, V, M=2), at the reference
initialisation and with randomised BatchNorm / graph_attn.  Run on the GPU box; output -> profiles/r2_parity_report.txt."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import continual_skeletons_b200 as cs  # noqa: E402
from oracle import step, weights  # noqa: E402

CASES = (
    (cs.CoStGcn, weights.cost_gcn_arch, "cost_gcn", "dummy_ntu", 25),
    (cs.CoStGcnMod, weights.cost_gcn_mod_arch, "cost_gcn_mod", "dummy_ntu", 25),
    (cs.CoAGcn, weights.coa_gcn_arch, "coa_gcn", "dummy_ntu", 25),
    (cs.CoSTr, weights.cos_tr_arch, "cos_tr", "dummy_kin", 18),
)
print("model          weights          path  chunk | max|logit|  max|d| vs reference blocks  vs step oracle  argmax equal | oracle vs reference blocks")
for cls, arch_fn, tag, dataset, V in CASES:
    gold = np.load(os.path.join(ROOT, "tests", "golden", tag + ".npz"))
    x = weights.make_input((2, 3, 300, V, 2), seed=11)
    for rnd in (False, True):
        arch = arch_fn()
        sd = weights.make_state_dict(arch, seed=8 if rnd else 7, randomize=rnd)
        ref = step.StepModel(sd, arch)
        with torch.no_grad():
            so = ref.forward_steps(x)
        want = torch.from_numpy(gold[f"{tag}_co_logits{'_rnd' if rnd else ''}"])
        for path, tc in (("auto", 1), ("auto", -1), ("simt", 1)):
            m = cls({"dataset_name": dataset, "kernel_path": path, "time_chunk": tc})
            m.load_state_dict(m.map_state_dict(sd), strict=True)
            out = m.forward_steps(x.cuda()).cpu()
            assert m.device_error() == 0
            print(f"{tag:13s}  {'randomised-BN' if rnd else 'reference-init':15s}  {path:4s}  {m._engine.time_chunk:5d} | {float(want.abs().max()):9.2f}  "
                  f"{float((out - want).abs().max()):25.2e}  {float((out - so).abs().max()):13.2e}  {str(bool(torch.equal(out.argmax(1), want.argmax(1)))):12s} | "
                  f"{float((so - want).abs().max()):.2e}")
print("\nchunk = frames per launch of forward_steps: 1 = frame by frame (64-channel blocks as one kernel per step), 300 = the whole clip module by module.")
print("North-star gate: max|d| <= 1e-3 at the reference initialisation.  The randomised-BN rows are a stress variant (logits up to ~100); they are held to 1e-3 * max|logit| / 16 in the tests.")
