"""Seeded weight fixtures in the *regular* (non-continual) ST-GCN state_dict key format.

Test infrastructure only.  The distributions follow the reference initialisers
(models/utils.py:9-24; models/base.py:235-257,299-300; models/st_gcn/st_gcn.py:44-46) but are drawn
from a numpy PCG64 stream so that fixtures are reproducible independent of torch's RNG.  The key
layout is the one SURVEY.md section 3.5 lists for ``StGcn.state_dict()``.
"""
import math
from collections import OrderedDict
from dataclasses import dataclass, field
from typing import List, Tuple

import numpy as np
import torch

from . import graphs


@dataclass
class BlockSpec:
    cin: int
    cout: int
    stride: int = 1
    residual: bool = True
    gconv: str = ""  # "" = the stack's ArchSpec.graph_conv; else "plain" / "adaptive" / "attention"

    @property
    def res_kind(self):
        """0 none / 1 identity / 2 strided 1x1 conv + BN (models/base.py:367-374)."""
        if not self.residual:
            return 0
        return 1 if (self.cin == self.cout and self.stride == 1) else 2


@dataclass
class ArchSpec:
    """Geometry of a stack: which blocks, which temporal padding, which skeleton."""

    blocks: List[BlockSpec]
    padding: int  # 4 ("equal", CoST-GCN) or 0 (CoST-GCN*)
    skeleton: str = "ntu"
    c_in: int = 3
    persons: int = 2
    classes: int = 60
    frames: int = 300
    head: bool = True  # data_bn + pool + fc around the blocks
    pool_size: int = -1
    pool_padding: int = -1
    block_names: List[str] = field(default_factory=list)
    # "plain": GraphConvolution (models/base.py:230-270); "adaptive": models/a_gcn/a_gcn.py:12-69;
    # "attention": GcnUnitAttention (models/s_tr/s_tr.py:303-476); a block's own BlockSpec.gconv overrides it
    graph_conv: str = "plain"

    def gconv_of(self, block):
        return block.gconv or self.graph_conv

    def __post_init__(self):
        if not self.block_names:
            self.block_names = [f"layers.layer{i + 1}." for i in range(len(self.blocks))]
        v = graphs.SKELETONS[self.skeleton][0]
        self.vertices = v
        # co.Sequential algebra, SURVEY.md section 3.3
        rf, cum, pad = 1, 1, 0
        for b in self.blocks:
            rf += 8 * cum
            pad += self.padding * cum
            cum *= b.stride
        self.receptive_field, self.stack_stride, self.stack_padding = rf, cum, pad
        if self.pool_size == -1:  # models/base.py:86-90
            self.pool_size = math.ceil((self.frames - rf + 2 * pad + 1) / cum)
        if self.pool_padding == -1:  # models/base.py:92-96
            self.pool_padding = self.pool_size - math.ceil((self.frames - rf + pad + 1) / cum)
        self.pool_padding = max(0, self.pool_padding)


def stgcn_blocks(c_in=3, strided=True) -> List[BlockSpec]:
    """The 10-block table of models/cost_gcn/cost_gcn.py:30-41 / cost_gcn_mod.py:29-40."""
    s = 2 if strided else 1
    return [
        BlockSpec(c_in, 64, 1, residual=False),
        BlockSpec(64, 64), BlockSpec(64, 64), BlockSpec(64, 64),
        BlockSpec(64, 128, s), BlockSpec(128, 128), BlockSpec(128, 128),
        BlockSpec(128, 256, s), BlockSpec(256, 256), BlockSpec(256, 256),
    ]


def cost_gcn_arch(skeleton="ntu", classes=60, **kw) -> ArchSpec:
    return ArchSpec(stgcn_blocks(strided=True), padding=4, skeleton=skeleton, classes=classes, **kw)


def cost_gcn_mod_arch(skeleton="ntu", classes=60, **kw) -> ArchSpec:
    return ArchSpec(stgcn_blocks(strided=False), padding=0, skeleton=skeleton, classes=classes, **kw)


def coa_gcn_arch(skeleton="ntu", classes=60, **kw) -> ArchSpec:
    """CoA-GCN: the CoST-GCN geometry with AdaptiveGraphConvolution (models/coa_gcn/coa_gcn.py:17-46)."""
    return ArchSpec(stgcn_blocks(strided=True), padding=4, skeleton=skeleton, classes=classes, graph_conv="adaptive", **kw)


def cos_tr_arch(skeleton="kinetics", classes=400, **kw) -> ArchSpec:
    """CoS-TR: CoST-GCN geometry, layers 4-10 with the spatial self-attention unit
    (models/cos_tr/cos_tr.py:24-41); BASELINE configs[3] quotes it on the 18-joint Kinetics skeleton."""
    blocks = stgcn_blocks(strided=True)
    for b in blocks[3:]:
        b.gconv = "attention"
    return ArchSpec(blocks, padding=4, skeleton=skeleton, classes=classes, **kw)


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))


def _conv(rng, sd, key, cout, cin, kt, bs):
    shape = (cout, cin, kt, 1)
    if bs == 1:  # kaiming normal, mode="fan_out"  (models/utils.py:12-13)
        std = math.sqrt(2.0 / (cout * kt))
    else:  # models/utils.py:15-19
        std = math.sqrt(2.0 / (cout * cin * kt * bs))
    sd[key + "weight"] = _t(rng.standard_normal(shape) * std)
    sd[key + "bias"] = torch.zeros(cout)


def _bn(sd, key, c, scale):
    sd[key + "weight"] = torch.full((c,), float(scale))
    sd[key + "bias"] = torch.zeros(c)
    sd[key + "running_mean"] = torch.zeros(c)
    sd[key + "running_var"] = torch.ones(c)
    sd[key + "num_batches_tracked"] = torch.zeros((), dtype=torch.long)


def _tail(rng, sd, name, b):
    """Temporal conv and block residual (models/base.py:291-300,367-374)."""
    _conv(rng, sd, name + "tcn.t_conv.", b.cout, b.cout, 9, bs=1)
    _bn(sd, name + "tcn.bn.", b.cout, 1)
    if b.res_kind == 2:
        _conv(rng, sd, name + "residual.t_conv.", b.cout, b.cin, 1, bs=1)
        _bn(sd, name + "residual.bn.", b.cout, 1)


def _torch_default_conv(rng, sd, key, cout, cin):
    """nn.Conv2d's own initialisation (kaiming_uniform with a = sqrt(5)): U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for
    weight and bias -- the attention unit never re-initialises its convs (models/s_tr/s_tr.py:82-101)."""
    bound = 1.0 / math.sqrt(cin)
    sd[key + "weight"] = _t(rng.uniform(-bound, bound, size=(cout, cin, 1, 1)))
    sd[key + "bias"] = _t(rng.uniform(-bound, bound, size=(cout,)))


def _attention_unit(rng, sd, g, b, A, V):
    """state_dict of GcnUnitAttention(only_attention=True) (models/s_tr/s_tr.py:303-415): data_bn over C*V
    features, bn, the parameter A, and SpatialAttention's qkv / output convs with dk = out/4, dv = out."""
    _bn(sd, g + "data_bn.", b.cin * V, 1)
    _bn(sd, g + "bn.", b.cout, 1)
    sd[g + "A"] = _t(A)
    dk, dv = int(b.cout * 0.25), b.cout
    _torch_default_conv(rng, sd, g + "attention_conv.qkv_conv.", 2 * dk + dv, b.cin)
    _torch_default_conv(rng, sd, g + "attention_conv.attn_out.", dv, dv)


def make_state_dict(arch: ArchSpec, seed: int, randomize: bool = False) -> "OrderedDict[str, torch.Tensor]":
    """Reference-style initial weights (``randomize=False``) or the randomised-BN variant that
    makes the adjacency branch visible."""
    rng = np.random.Generator(np.random.PCG64(seed))
    A = graphs.adjacency(arch.skeleton)
    V = arch.vertices
    sd = OrderedDict()
    if arch.head:
        _bn(sd, "data_bn.", arch.persons * arch.c_in * V, 1)
    for name, b in zip(arch.block_names, arch.blocks):
        g = name + "gcn."
        kind = arch.gconv_of(b)
        if kind == "attention":
            _attention_unit(rng, sd, g, b, A, V)
            _tail(rng, sd, name, b)
            continue
        sd[g + "graph_attn"] = torch.ones(3, V, V)
        sd[g + "A"] = _t(A)
        for i in range(3):
            _conv(rng, sd, g + f"g_conv.{i}.", b.cout, b.cin, 1, bs=3)
        if kind == "adaptive":  # a_gcn.py:14,25-31 (coff_embedding = 4)
            inter_c = b.cout // 4
            for i in range(3):
                _conv(rng, sd, g + f"a_conv.{i}.", inter_c, b.cin, 1, bs=1)
                _conv(rng, sd, g + f"b_conv.{i}.", inter_c, b.cin, 1, bs=1)
        if b.cin != b.cout:
            _conv(rng, sd, g + "gcn_residual.0.", b.cout, b.cin, 1, bs=1)
            _bn(sd, g + "gcn_residual.1.", b.cout, 1)
        _bn(sd, g + "bn.", b.cout, 1e-6)
        _tail(rng, sd, name, b)
    if arch.head:
        c_last = arch.blocks[-1].cout
        sd["fc.weight"] = _t(rng.standard_normal((arch.classes, c_last)) * math.sqrt(2.0 / arch.classes))
        bound = 1.0 / math.sqrt(c_last)
        sd["fc.bias"] = _t(rng.uniform(-bound, bound, size=(arch.classes,)))
    if randomize:
        for k in list(sd.keys()):
            v = sd[k]
            if k.endswith("num_batches_tracked") or k.endswith(".A"):
                continue
            is_bn = (k[: k.rfind(".")] + ".running_var") in sd
            if k.endswith("graph_attn"):
                if arch.graph_conv == "adaptive":
                    # dense additive term (a_gcn.py:50): keep it small so that ten layers of it do not blow the
                    # activations up to where every attention softmax saturates into a hard arg-max
                    sd[k] = _t(rng.uniform(-0.04, 0.08, size=tuple(v.shape)))
                else:
                    sd[k] = _t(rng.uniform(0.5, 1.5, size=tuple(v.shape)))
            elif k.endswith("running_var"):
                sd[k] = _t(rng.uniform(0.5, 1.5, size=tuple(v.shape)))
            elif k.endswith("running_mean"):
                sd[k] = _t(rng.uniform(-0.2, 0.2, size=tuple(v.shape)))
            elif is_bn and k.endswith("weight"):
                sd[k] = _t(rng.uniform(0.5, 1.5, size=tuple(v.shape)))
            elif is_bn and k.endswith("bias"):
                sd[k] = _t(rng.uniform(-0.2, 0.2, size=tuple(v.shape)))
            elif k.endswith("bias") and not k.startswith("fc."):
                sd[k] = _t(rng.uniform(-0.1, 0.1, size=tuple(v.shape)))
    return sd


def make_input(shape: Tuple[int, ...], seed: int) -> torch.Tensor:
    """U[0,1) frames, the DummyDataset distribution (datasets/datasets.py:301)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    return _t(rng.random(size=shape))
