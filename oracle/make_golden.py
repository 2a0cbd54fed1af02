"""Generate tests/golden/*.npz by running the REFERENCE's own blocks (imported from
/root/reference through oracle/ref_shim.py) on seeded weights and inputs.

Run in the build container only:  ``python -m oracle.make_golden``.
The fixtures are committed; the GPU box never needs /root/reference.

What is reference code here: ``SpatioTemporalBlock`` (and through it ``GraphConvolution`` /
``TemporalConvolution``), ``graph.A``.  What is glue restated from models/st_gcn/st_gcn.py:48-65
(the class itself cannot be imported without ``ride``): the data_bn reshapes, the mean and fc.
"""
import os
import sys

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import ref_shim, weights  # noqa: E402
from oracle.weights import BlockSpec  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# name, cin, cout, stride, residual, temporal_padding  (cases of tests/test_cost_gcn.py:71-271 and
# tests/test_st_gcn_mod.py:11-54, plus wide-channel cases that exercise the tensor-core tiles)
BLOCK_CASES = [
    ("nores_p4", 4, 4, 1, False, 4),
    ("idres_p4", 4, 4, 1, True, 4),
    ("idres_p0", 4, 4, 1, True, 0),
    ("convres_p4", 2, 4, 1, True, 4),
    ("convres_s2_p4", 2, 4, 2, True, 4),
    ("convres_p0", 2, 4, 1, True, 0),
    ("wide_idres_p4", 64, 64, 1, True, 4),
    ("wide_convres_s2_p4", 64, 128, 2, True, 4),
    ("wide_convres_p0", 128, 256, 1, True, 0),
]
BLOCK_T, BLOCK_B = 24, 2

# the same block with AdaptiveGraphConvolution evaluated one frame at a time (CoAGcn's step semantics:
# models/coa_gcn/coa_gcn.py:11-14 wraps the module in co.forward_stepping, i.e. T = 1 per call)
ADAPTIVE_BLOCK_CASES = [
    ("a_nores_p4", 4, 4, 1, False, 4),
    ("a_idres_p4", 4, 4, 1, True, 4),
    ("a_convres_s2_p4", 2, 4, 2, True, 4),
    ("a_first_p4", 3, 64, 1, False, 4),
    ("a_wide_idres_p4", 64, 64, 1, True, 4),
    ("a_wide_convres_s2_p4", 64, 128, 2, True, 4),
    ("a_wide_convres_p4", 128, 256, 1, True, 4),
    ("a_wide_idres256_p4", 256, 256, 1, True, 4),
]


# blocks whose graph conv is the spatial self-attention unit of CoS-TR (models/s_tr/s_tr.py:303-476; it works on one
# frame at a time by construction, so clip and step semantics coincide).  out_channels must be a multiple of 32
# (dk = out/4 split over 8 heads).  name, cin, cout, stride, residual, temporal_padding, skeleton
ATTENTION_BLOCK_CASES = [
    ("s_idres_p4", 32, 32, 1, True, 4, "ntu"),
    ("s_convres_s2_p4", 8, 32, 2, True, 4, "ntu"),
    ("s_wide_idres_p4", 64, 64, 1, True, 4, "ntu"),
    ("s_wide_convres_s2_p4", 64, 128, 2, True, 4, "kinetics"),
    ("s_wide_idres256_p4", 256, 256, 1, True, 4, "kinetics"),
]


def attention_unit(ref, vertices):
    """GraphConv factory as models/s_tr/s_tr.py:499-502 / models/cos_tr/cos_tr.py:25-28 build it."""

    def make(in_channels, out_channels, A):
        return ref.GcnUnitAttention(in_channels, out_channels, A, num_point=vertices)

    return make


def per_frame_adaptive(ref):
    """The reference's AdaptiveGraphConvolution applied frame by frame: what forward_step computes, laid
    out as a clip so that the reference's SpatioTemporalBlock can run the temporal part."""

    class PerFrameAdaptiveGraphConvolution(ref.AdaptiveGraphConvolution):
        def forward(self, x):
            fwd = super().forward
            return torch.cat([fwd(x[:, :, t: t + 1]) for t in range(x.shape[2])], dim=2)

    return PerFrameAdaptiveGraphConvolution


class RefStack(nn.Module):
    """Reference SpatioTemporalBlocks wired like StGcn / StGcnMod (st_gcn.py:27-46)."""

    def __init__(self, ref, arch, A):
        super().__init__()
        if arch.head:
            self.data_bn = nn.BatchNorm1d(arch.persons * arch.c_in * arch.vertices)
        tp = -1 if arch.padding == 4 else arch.padding

        def kw(b):
            kind = arch.gconv_of(b)
            if kind == "adaptive":
                return {"GraphConv": per_frame_adaptive(ref)}
            if kind == "attention":
                return {"GraphConv": attention_unit(ref, arch.vertices)}
            return {}

        self.layers = nn.ModuleDict(
            {
                n.split(".")[-2]: ref.SpatioTemporalBlock(b.cin, b.cout, A, stride=b.stride, residual=b.residual, temporal_padding=tp, **kw(b))
                for n, b in zip(arch.block_names, arch.blocks)
            }
        )
        if arch.head:
            self.fc = nn.Linear(arch.blocks[-1].cout, arch.classes)
        self.arch = arch

    def features(self, x, collect=None):
        N, C, T, V, M = x.size()
        x = x.permute(0, 4, 3, 1, 2).contiguous().view(N, M * V * C, T)
        x = self.data_bn(x)
        x = x.view(N, M, V, C, T).permute(0, 1, 3, 4, 2).contiguous().view(N * M, C, T, V)
        for k in self.layers:
            x = self.layers[k](x)
            if collect is not None:
                collect.append(x)
        return x

    def forward(self, x):
        N, M = x.shape[0], x.shape[4]
        y = self.features(x)
        y = y.view(N, M, y.size(1), -1).mean(3).mean(1)
        return self.fc(y)


def block_fixtures(ref):
    out = {}
    A = ref.ntu_A
    for idx, (name, cin, cout, stride, residual, pad) in enumerate(BLOCK_CASES):
        for rnd in (False, True):
            arch = weights.ArchSpec([BlockSpec(cin, cout, stride, residual)], padding=pad, head=False, block_names=[""])
            sd = weights.make_state_dict(arch, seed=1000 + idx, randomize=rnd)
            blk = ref.SpatioTemporalBlock(cin, cout, A, stride=stride, residual=residual, temporal_padding=pad)
            blk.load_state_dict(sd, strict=True)
            blk.eval()
            batch = 1 if name.startswith("wide") else BLOCK_B  # keep the fixtures small
            x = weights.make_input((batch, cin, BLOCK_T, 25), seed=2000 + idx)
            with torch.no_grad():
                y = blk(x)
            out[f"{name}{'_rnd' if rnd else ''}"] = y.numpy()
    return out


def adaptive_block_fixtures(ref):
    out = {}
    A = ref.ntu_A
    gc = per_frame_adaptive(ref)
    for idx, (name, cin, cout, stride, residual, pad) in enumerate(ADAPTIVE_BLOCK_CASES):
        for rnd in (False, True):
            arch = weights.ArchSpec([BlockSpec(cin, cout, stride, residual)], padding=pad, head=False, block_names=[""], graph_conv="adaptive")
            sd = weights.make_state_dict(arch, seed=3000 + idx, randomize=rnd)
            blk = ref.SpatioTemporalBlock(cin, cout, A, stride=stride, residual=residual, temporal_padding=pad, GraphConv=gc)
            blk.load_state_dict(sd, strict=True)
            blk.eval()
            batch = 1 if "wide" in name else BLOCK_B
            frames = 14 if "wide" in name else BLOCK_T  # keep the fixture small
            x = weights.make_input((batch, cin, frames, 25), seed=4000 + idx)
            with torch.no_grad():
                y = blk(x)
                g = blk.gcn(x[:, :, :2])  # graph conv alone on two frames
            out[f"{name}{'_rnd' if rnd else ''}"] = y.numpy()
            out[f"{name}{'_rnd' if rnd else ''}_gcn"] = g.numpy()
    return out


def attention_block_fixtures(ref):
    out = {}
    for idx, (name, cin, cout, stride, residual, pad, skel) in enumerate(ATTENTION_BLOCK_CASES):
        A = ref.ntu_A if skel == "ntu" else ref.kinetics_A
        V = A.shape[-1]
        for rnd in (False, True):
            arch = weights.ArchSpec([BlockSpec(cin, cout, stride, residual, gconv="attention")], padding=pad, head=False,
                                    block_names=[""], skeleton=skel)
            sd = weights.make_state_dict(arch, seed=5000 + idx, randomize=rnd)
            blk = ref.SpatioTemporalBlock(cin, cout, A, stride=stride, residual=residual, temporal_padding=pad,
                                          GraphConv=attention_unit(ref, V))
            blk.load_state_dict(sd, strict=True)
            blk.eval()
            batch = 1 if "wide" in name else BLOCK_B
            frames = 14 if "wide" in name else BLOCK_T
            x = weights.make_input((batch, cin, frames, V), seed=6000 + idx)
            with torch.no_grad():
                y = blk(x)
                g = blk.gcn(x[:, :, :2])
            out[f"{name}{'_rnd' if rnd else ''}"] = y.numpy()
            out[f"{name}{'_rnd' if rnd else ''}_gcn"] = g.numpy()
    return out


def model_fixtures(ref, arch_fn, tag, n=2):
    out = {}
    for rnd in (False, True):
        arch = arch_fn()
        sd = weights.make_state_dict(arch, seed=7 if not rnd else 8, randomize=rnd)
        net = RefStack(ref, arch, ref.ntu_A if arch.skeleton == "ntu" else ref.kinetics_A)
        missing = net.load_state_dict(sd, strict=True)
        assert not missing.missing_keys and not missing.unexpected_keys
        net.eval()
        x = weights.make_input((n, arch.c_in, arch.frames, arch.vertices, arch.persons), seed=11)
        with torch.no_grad():
            per_block = []
            feats = net.features(x, per_block)
            _, c, t, v = feats.shape
            pooled = feats.view(n, arch.persons, c, t, v).mean(4).mean(1)  # base.py:84
            reg_logits = net(x)
            win = F.avg_pool1d(pooled, arch.pool_size, stride=1, padding=arch.pool_padding)
            co_logits = (torch.einsum("kc,nct->nkt", net.fc.weight, win) + net.fc.bias[None, :, None])[:, :, 0]
        sfx = "_rnd" if rnd else ""
        out[f"{tag}_reg_logits{sfx}"] = reg_logits.numpy()
        out[f"{tag}_co_logits{sfx}"] = co_logits.numpy()
        out[f"{tag}_pooled0{sfx}"] = pooled[0].numpy()
        for li, y in enumerate(per_block):  # one time slice of skeleton 0 per block
            out[f"{tag}_block{li + 1}_mid{sfx}"] = y[0, :, y.shape[2] // 2].numpy()
    return out


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    ref = ref_shim.load()
    os.makedirs(OUT, exist_ok=True)
    only = set(sys.argv[1:])  # e.g. ``python -m oracle.make_golden coa_gcn`` regenerates just that file

    def save(name, make):
        if not only or name in only:
            np.savez_compressed(os.path.join(OUT, name + ".npz"), **make())

    save("adjacency", lambda: dict(ntu=ref.ntu_A, kinetics=ref.kinetics_A))
    save("blocks", lambda: block_fixtures(ref))
    save("cost_gcn", lambda: model_fixtures(ref, weights.cost_gcn_arch, "cost_gcn"))
    save("cost_gcn_mod", lambda: model_fixtures(ref, weights.cost_gcn_mod_arch, "cost_gcn_mod"))
    save("coa_blocks", lambda: adaptive_block_fixtures(ref))
    save("coa_gcn", lambda: model_fixtures(ref, weights.coa_gcn_arch, "coa_gcn"))
    save("cos_blocks", lambda: attention_block_fixtures(ref))
    save("cos_tr", lambda: model_fixtures(ref, weights.cos_tr_arch, "cos_tr"))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
