"""CPU tests of the oracle: restated reference math vs the committed golden fixtures (made from
the reference's own blocks by oracle/make_golden.py) and the continual step oracle vs the clip
oracle through the relations the reference's tests pin (tests/test_cost_gcn.py:37-362,
tests/test_st_gcn_mod.py:11-90 in the reference tree)."""
import numpy as np
import pytest
import torch

from oracle import graphs, ref_shim, regular, step, weights
from oracle.make_golden import ADAPTIVE_BLOCK_CASES, ATTENTION_BLOCK_CASES, BLOCK_B, BLOCK_CASES, BLOCK_T
from oracle.weights import ArchSpec, BlockSpec


def _block_case(idx, rnd):
    name, cin, cout, stride, residual, pad = BLOCK_CASES[idx]
    arch = ArchSpec([BlockSpec(cin, cout, stride, residual)], padding=pad, head=False, block_names=[""])
    sd = weights.make_state_dict(arch, seed=1000 + idx, randomize=rnd)
    batch = 1 if name.startswith("wide") else BLOCK_B
    x = weights.make_input((batch, cin, BLOCK_T, 25), seed=2000 + idx)
    return name + ("_rnd" if rnd else ""), arch, sd, x


def test_adjacency_bit_exact(golden):
    for name in ("ntu", "kinetics"):
        a = graphs.adjacency(name)
        assert a.dtype == np.float64
        assert np.array_equal(a, golden["adjacency"][name])
    nz = [(graphs.adjacency("ntu")[i] != 0).sum() for i in range(3)]
    assert nz == [25, 24, 24]


@pytest.mark.parametrize("idx", range(len(BLOCK_CASES)))
@pytest.mark.parametrize("rnd", [False, True])
def test_regular_block_vs_golden(golden, idx, rnd):
    key, arch, sd, x = _block_case(idx, rnd)
    y = regular.st_block(x, sd, "", arch.blocks[0], arch.padding)
    ref = torch.from_numpy(golden["blocks"][key])
    assert y.shape == ref.shape
    assert torch.allclose(y, ref, atol=1e-6, rtol=1e-6)


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree only exists in the build container")
@pytest.mark.parametrize("idx", [1, 4, 5])
def test_regular_block_vs_live_reference(idx):
    ref = ref_shim.load()
    key, arch, sd, x = _block_case(idx, True)
    b = arch.blocks[0]
    blk = ref.SpatioTemporalBlock(b.cin, b.cout, ref.ntu_A, stride=b.stride, residual=b.residual, temporal_padding=arch.padding)
    blk.load_state_dict(sd, strict=True)
    blk.eval()
    with torch.no_grad():
        assert torch.allclose(regular.st_block(x, sd, "", b, arch.padding), blk(x), atol=1e-6)
    assert np.array_equal(ref.ntu_A, graphs.adjacency("ntu"))


@pytest.mark.parametrize("idx", range(6))
@pytest.mark.parametrize("rnd", [False, True])
def test_step_block_vs_regular(golden, idx, rnd):
    """Per-frame output at index t + (8 - p) equals clip output t (test_cost_gcn.py:114-126);
    strided blocks emit output j at step 2 j + 4."""
    key, arch, sd, x = _block_case(idx, rnd)
    spec, p = arch.blocks[0], arch.padding
    target = torch.from_numpy(golden["blocks"][key])
    blk = step.StepBlock(sd, "", spec, p)
    outs = [blk.step(x[:, :, t]) for t in range(x.shape[2])]
    first = 8 - p
    for t, o in enumerate(outs):
        due = t >= first and (t - first) % spec.stride == 0
        assert (o is not None) == due
    emitted = [o for o in outs if o is not None]
    # clip output index j; with p=4 the clip's own first output j=0 sees 4 start zeros like the ring
    for j, o in enumerate(emitted):
        assert torch.allclose(o, target[:, :, j], atol=1e-6), (key, j)


def test_step_stack_prefix_equals_clip():
    """3-block stack (no-res, id-res, strided conv-res): forward_steps(pad_end=False) equals the
    regular clip output minus the outputs that would need end padding
    (test_cost_gcn.py:274-326 with the :218-220 / :264-266 prefix relation)."""
    arch = ArchSpec([BlockSpec(3, 3, 1, False), BlockSpec(3, 3), BlockSpec(3, 4, 2)], padding=4, head=False,
                    block_names=["0.", "1.", "2."])
    assert (arch.receptive_field, arch.stack_stride, arch.stack_padding) == (25, 2, 12)
    sd = weights.make_state_dict(arch, seed=5, randomize=True)
    x = weights.make_input((2, 3, 40, 25), seed=6)
    target = regular.stack_features(x, sd, arch)
    m = step.StepModel(sd, arch)
    out = m.forward_steps(x)
    n_out = out.shape[2]
    assert n_out == target.shape[2] - arch.stack_padding // arch.stack_stride
    assert torch.allclose(out, target[:, :, :n_out], atol=1e-6)
    with pytest.raises(NotImplementedError):
        m.forward_steps(x, pad_end=True)


@pytest.mark.parametrize("rnd", [False, True])
def test_step_model_cost_gcn_mod(golden, rnd):
    """CoStGcnMod: forward_steps == clip forward == regular StGcnMod (test_st_gcn_mod.py:66-90)."""
    g, sfx = golden["cost_gcn_mod"], "_rnd" if rnd else ""
    arch = weights.cost_gcn_mod_arch()
    assert (arch.receptive_field, arch.stack_stride, arch.stack_padding, arch.pool_size, arch.pool_padding) == (81, 1, 0, 220, 0)
    sd = weights.make_state_dict(arch, seed=8 if rnd else 7, randomize=rnd)
    x = weights.make_input((2, 3, 300, 25, 2), seed=11)
    m = step.StepModel(sd, arch)
    with torch.no_grad():
        out = m.forward_steps(x)
    assert out.shape == (2, 60)
    emits = [i for i, f in enumerate(m.trace) if f[-1]]
    assert emits == [299]
    scale = max(1.0, float(np.abs(g["cost_gcn_mod_reg_logits" + sfx]).max()) / 16)
    assert torch.allclose(out, torch.from_numpy(g["cost_gcn_mod_reg_logits" + sfx]), atol=5e-4 * scale)
    assert torch.allclose(out, torch.from_numpy(g["cost_gcn_mod_co_logits" + sfx]), atol=5e-4 * scale)


@pytest.mark.parametrize("rnd", [False, True])
def test_step_model_cost_gcn(golden, rnd):
    """CoStGcn: one emission at frame 296, equal to the clip-mode continual forward and top-3
    consistent with regular StGcn (test_cost_gcn.py:329-362)."""
    g, sfx = golden["cost_gcn"], "_rnd" if rnd else ""
    arch = weights.cost_gcn_arch()
    assert (arch.receptive_field, arch.stack_stride, arch.stack_padding, arch.pool_size, arch.pool_padding) == (153, 4, 76, 75, 19)
    sd = weights.make_state_dict(arch, seed=8 if rnd else 7, randomize=rnd)
    x = weights.make_input((2, 3, 300, 25, 2), seed=11)
    m = step.StepModel(sd, arch)
    with torch.no_grad():
        out = m.forward_steps(x)
    assert out.shape == (2, 60)
    assert [i for i, f in enumerate(m.trace) if f[-1]] == [296]
    # layer 10 first emits at frame 76, then every 4th frame
    l10 = [i for i, f in enumerate(m.trace) if f[9]]
    assert l10[0] == 76 and all(b - a == 4 for a, b in zip(l10, l10[1:]))
    co = torch.from_numpy(g["cost_gcn_co_logits" + sfx])
    reg = torch.from_numpy(g["cost_gcn_reg_logits" + sfx])
    assert torch.allclose(out, co, rtol=1e-4, atol=1e-4)
    # The 56/75 pooling makes the logits differ from StGcn's; the reference pins the top-3 on its
    # own random instance (test_cost_gcn.py:352-362).  Near ties may swap ranks 2/3, so compare
    # the argmax and the top-3 as a set.
    assert torch.equal(out.argmax(1), reg.argmax(1))
    for a, b in zip(torch.topk(out, 3).indices.tolist(), torch.topk(reg, 3).indices.tolist()):
        assert set(a) == set(b)


def test_regular_model_vs_golden(golden):
    for tag, fn in (("cost_gcn", weights.cost_gcn_arch), ("cost_gcn_mod", weights.cost_gcn_mod_arch)):
        arch = fn()
        sd = weights.make_state_dict(arch, seed=8, randomize=True)
        x = weights.make_input((2, 3, 300, 25, 2), seed=11)
        g = golden[tag]
        with torch.no_grad():
            blocks = []
            logits = regular.stgcn_forward(x, sd, arch, blocks)
            co = regular.co_clip_forward(x, sd, arch)
        scale = max(1.0, float(np.abs(g[tag + "_reg_logits_rnd"]).max()) / 16)
        assert torch.allclose(logits, torch.from_numpy(g[tag + "_reg_logits_rnd"]), atol=1e-4 * scale)
        assert torch.allclose(co, torch.from_numpy(g[tag + "_co_logits_rnd"]), atol=1e-4 * scale)
        for li, y in enumerate(blocks):
            ref = torch.from_numpy(g[f"{tag}_block{li + 1}_mid_rnd"])
            got = y[0, :, y.shape[2] // 2]
            assert torch.allclose(got, ref, atol=1e-5 * max(1.0, float(ref.abs().max()))), li


def test_step_oracle_kinetics_graph():
    """The step oracle on the 18-joint Kinetics graph (strided conv-residual block): prefix relation."""
    arch = ArchSpec([BlockSpec(4, 4, 1, False), BlockSpec(4, 8, 2)], padding=4, skeleton="kinetics", head=False,
                    block_names=["0.", "1."])
    sd = weights.make_state_dict(arch, seed=41, randomize=True)
    x = weights.make_input((2, 4, 30, 18), seed=42)
    target = regular.stack_features(x, sd, arch)
    out = step.StepModel(sd, arch).forward_steps(x)
    n = out.shape[2]
    assert n == target.shape[2] - arch.stack_padding // arch.stack_stride
    assert torch.allclose(out, target[:, :, :n], atol=1e-6)


# ---------------------------------------------------------------------------------------------
# CoA-GCN: AdaptiveGraphConvolution evaluated one frame at a time
# ---------------------------------------------------------------------------------------------
def _adaptive_case(idx, rnd):
    name, cin, cout, stride, residual, pad = ADAPTIVE_BLOCK_CASES[idx]
    arch = ArchSpec([BlockSpec(cin, cout, stride, residual)], padding=pad, head=False, block_names=[""], graph_conv="adaptive")
    sd = weights.make_state_dict(arch, seed=3000 + idx, randomize=rnd)
    wide = "wide" in name
    x = weights.make_input((1 if wide else BLOCK_B, cin, 14 if wide else BLOCK_T, 25), seed=4000 + idx)
    return name + ("_rnd" if rnd else ""), arch, sd, x


@pytest.mark.parametrize("idx", range(len(ADAPTIVE_BLOCK_CASES)))
@pytest.mark.parametrize("rnd", [False, True])
def test_adaptive_block_vs_golden(golden, idx, rnd):
    """Clip-layout restatement with per-frame attention, and the step oracle, against the reference's own
    AdaptiveGraphConvolution + SpatioTemporalBlock (fixtures by oracle/make_golden.py)."""
    key, arch, sd, x = _adaptive_case(idx, rnd)
    spec, p = arch.blocks[0], arch.padding
    ref = torch.from_numpy(golden["coa_blocks"][key])
    scale = max(1.0, float(ref.abs().max()))
    y = regular.st_block(x, sd, "", spec, p, per_frame=True)
    assert y.shape == ref.shape and torch.allclose(y, ref, atol=2e-6 * scale)
    g = regular.graph_conv(x[:, :, :2], sd, "gcn.", per_frame=True)
    assert torch.allclose(g, torch.from_numpy(golden["coa_blocks"][key + "_gcn"]), atol=2e-6 * scale)
    blk = step.StepBlock(sd, "", spec, p)
    emitted = [o for o in (blk.step(x[:, :, t]) for t in range(x.shape[2])) if o is not None]
    assert len(emitted) == (x.shape[2] - (8 - p) + spec.stride - 1) // spec.stride
    for j, o in enumerate(emitted):
        assert torch.allclose(o, ref[:, :, j], atol=2e-6 * scale), (key, j)


def test_adaptive_attention_spans_the_clip():
    """The reason CoA-GCN is not equivalent to A-GCN (models/coa_gcn/coa_gcn.py:29): over a clip the reference
    normalises one attention map over all frames, per step it sees one frame."""
    key, arch, sd, x = _adaptive_case(1, True)
    whole = regular.graph_conv(x, sd, "gcn.")
    framewise = regular.graph_conv(x, sd, "gcn.", per_frame=True)
    assert whole.shape == framewise.shape and not torch.allclose(whole, framewise, atol=1e-3)


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree only exists in the build container")
def test_adaptive_graph_conv_vs_live_reference():
    ref = ref_shim.load()
    for idx in (1, 5):
        key, arch, sd, x = _adaptive_case(idx, True)
        b = arch.blocks[0]
        gc = ref.AdaptiveGraphConvolution(b.cin, b.cout, ref.ntu_A)
        gc.load_state_dict({k[len("gcn."):]: v for k, v in sd.items() if k.startswith("gcn.")}, strict=True)
        gc.eval()
        with torch.no_grad():
            assert torch.allclose(regular.graph_conv(x, sd, "gcn."), gc(x), atol=1e-6)  # whole-clip attention
            assert torch.allclose(regular.graph_conv(x[:, :, 3:4], sd, "gcn."), gc(x[:, :, 3:4]), atol=1e-6)


@pytest.mark.parametrize("rnd", [False, True])
def test_step_model_coa_gcn(golden, rnd):
    """CoAGcn steps: same schedule as CoStGcn (first logits at frame 296), logits equal to the reference blocks
    run with per-frame attention."""
    g, sfx = golden["coa_gcn"], "_rnd" if rnd else ""
    arch = weights.coa_gcn_arch()
    assert (arch.receptive_field, arch.stack_stride, arch.stack_padding, arch.pool_size, arch.pool_padding) == (153, 4, 76, 75, 19)
    sd = weights.make_state_dict(arch, seed=8 if rnd else 7, randomize=rnd)
    x = weights.make_input((2, 3, 300, 25, 2), seed=11)
    m = step.StepModel(sd, arch)
    with torch.no_grad():
        out = m.forward_steps(x)
    assert out.shape == (2, 60)
    assert [i for i, f in enumerate(m.trace) if f[-1]] == [296]
    co = torch.from_numpy(g["coa_gcn_co_logits" + sfx])
    scale = max(1.0, float(co.abs().max()) / 16)
    assert torch.allclose(out, co, rtol=1e-4, atol=1e-4 * scale)


# ---------------------------------------------------------------------------------------------
# CoS-TR: spatial self-attention unit in layers 4-10
# ---------------------------------------------------------------------------------------------
def _attention_case(idx, rnd):
    name, cin, cout, stride, residual, pad, skel = ATTENTION_BLOCK_CASES[idx]
    arch = ArchSpec([BlockSpec(cin, cout, stride, residual, gconv="attention")], padding=pad, head=False, block_names=[""], skeleton=skel)
    sd = weights.make_state_dict(arch, seed=5000 + idx, randomize=rnd)
    wide = "wide" in name
    x = weights.make_input((1 if wide else BLOCK_B, cin, 14 if wide else BLOCK_T, arch.vertices), seed=6000 + idx)
    return name + ("_rnd" if rnd else ""), arch, sd, x


@pytest.mark.parametrize("idx", range(len(ATTENTION_BLOCK_CASES)))
@pytest.mark.parametrize("rnd", [False, True])
def test_attention_block_vs_golden(golden, idx, rnd):
    """Restated GcnUnitAttention block (clip form and stepped) against fixtures made by the reference's own
    GcnUnitAttention + SpatioTemporalBlock on the NTU and Kinetics graphs."""
    key, arch, sd, x = _attention_case(idx, rnd)
    spec, p = arch.blocks[0], arch.padding
    ref = torch.from_numpy(golden["cos_blocks"][key])
    scale = max(1.0, float(ref.abs().max()))
    y = regular.st_block(x, sd, "", spec, p)
    assert y.shape == ref.shape and torch.allclose(y, ref, atol=2e-6 * scale)
    g = regular.graph_conv(x[:, :, :2], sd, "gcn.")
    assert torch.allclose(g, torch.from_numpy(golden["cos_blocks"][key + "_gcn"]), atol=2e-6 * scale)
    blk = step.StepBlock(sd, "", spec, p)
    emitted = [o for o in (blk.step(x[:, :, t]) for t in range(x.shape[2])) if o is not None]
    assert len(emitted) == (x.shape[2] - (8 - p) + spec.stride - 1) // spec.stride
    for j, o in enumerate(emitted):
        assert torch.allclose(o, ref[:, :, j], atol=2e-6 * scale), (key, j)


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree only exists in the build container")
def test_attention_unit_vs_live_reference():
    ref = ref_shim.load()
    for idx in (0, 3):
        key, arch, sd, x = _attention_case(idx, True)
        b = arch.blocks[0]
        A = ref.ntu_A if arch.skeleton == "ntu" else ref.kinetics_A
        unit = ref.GcnUnitAttention(b.cin, b.cout, A, num_point=arch.vertices)
        unit.load_state_dict({k[len("gcn."):]: v for k, v in sd.items() if k.startswith("gcn.")}, strict=True)
        unit.eval()
        with torch.no_grad():
            assert torch.allclose(regular.graph_conv(x, sd, "gcn."), unit(x), atol=1e-6)


@pytest.mark.parametrize("rnd", [False, True])
def test_step_model_cos_tr(golden, rnd):
    """CoSTr on the Kinetics skeleton (V = 18, 400 classes): CoST-GCN's schedule, logits equal to the reference blocks."""
    g, sfx = golden["cos_tr"], "_rnd" if rnd else ""
    arch = weights.cos_tr_arch()
    assert (arch.vertices, arch.receptive_field, arch.stack_stride, arch.stack_padding, arch.pool_size, arch.pool_padding) == (18, 153, 4, 76, 75, 19)
    assert [arch.gconv_of(b) for b in arch.blocks] == ["plain"] * 3 + ["attention"] * 7
    sd = weights.make_state_dict(arch, seed=8 if rnd else 7, randomize=rnd)
    x = weights.make_input((2, 3, 300, 18, 2), seed=11)
    m = step.StepModel(sd, arch)
    with torch.no_grad():
        out = m.forward_steps(x)
    assert out.shape == (2, 400)
    assert [i for i, f in enumerate(m.trace) if f[-1]] == [296]
    co = torch.from_numpy(g["cos_tr_co_logits" + sfx])
    assert torch.allclose(out, co, rtol=1e-4, atol=1e-4 * max(1.0, float(co.abs().max()) / 16))
