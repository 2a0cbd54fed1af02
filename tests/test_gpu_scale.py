"""GPU parity at scale (``-m gpu``): DISTINCT random streams at a stream count with more token tiles than
the B200 has SMs (persistent-loop wrap, odd tile count, ragged last tile, reversed tile walk, CTA-pair
phantom tile), compared stream by stream with the CPU oracle; late pooled emissions against the oracle
on a non-constant input; CoA-GCN on the NTU RGB+D 120 geometry BASELINE configs[2] names; and the
sharded-vs-single-GPU bit-equality of SURVEY section 8e over NCCL (needs two devices).

Every parity figure is printed UNSCALED (max|d|, max|logit|).  The north-star gate -- max |dlogit| <= 1e-3
at the reference's random init -- is asserted in absolute terms for all four models; the randomised-BN
stress variants (logits up to ~100) are asserted relative to max|logit| / 16 and their absolute figure is
printed and recorded in profiles/r2_parity_report.txt.
"""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import continual_skeletons_b200 as cs
from oracle import regular, step, weights

pytestmark = pytest.mark.gpu

DEV = "cuda:0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

MODELS = {
    # tag: (class, arch factory, dataset, V, per-frame graph conv in the clip oracle)
    "cost_gcn": (cs.CoStGcn, weights.cost_gcn_arch, "dummy_ntu", 25, False),
    "cost_gcn_mod": (cs.CoStGcnMod, weights.cost_gcn_mod_arch, "dummy_ntu", 25, False),
    "coa_gcn": (cs.CoAGcn, weights.coa_gcn_arch, "dummy_ntu", 25, True),
    "cos_tr": (cs.CoSTr, weights.cos_tr_arch, "dummy_kin", 18, True),
}


def _rel_err(got, want):
    return float((got - want).abs().max()) / max(1.0, float(want.abs().max()))


def _build(tag, rnd, **hp):
    cls, arch_fn, dataset, V, per_frame = MODELS[tag]
    arch = arch_fn(**{k: v for k, v in hp.items() if k in ("pool_size", "pool_padding", "classes")})
    sd = weights.make_state_dict(arch, seed=8 if rnd else 7, randomize=rnd)
    m = cls({"dataset_name": hp.get("dataset_name", dataset), "pool_size": hp.get("pool_size", -1),
             "pool_padding": hp.get("pool_padding", -1)})
    m.load_state_dict(m.map_state_dict(sd), strict=True)
    return arch, sd, m, V, per_frame


@pytest.mark.parametrize("rnd", [False, True])
@pytest.mark.parametrize("tag", list(MODELS))
def test_distinct_streams_more_tiles_than_sms(tag, rnd):
    """1037 distinct streams (415 token tiles for V = 25, 297 for V = 18: odd, > 2 x 148 SMs, ragged tail).  A random
    subset of 64 streams -- always including the first and last stream and both sides of the persistent loop's wrap --
    is pushed through the CPU step oracle; logits, the emission schedule and the layer-5 / 8 / 10 block outputs of
    exactly those streams must agree.  A short pooling window (8, no padding) brings the first logits forward so
    that the oracle finishes in seconds; the 300-frame window is covered by the 2-stream north-star tests."""
    arch, sd, m, V, per_frame = _build(tag, rnd, pool_size=8, pool_padding=0)
    N = 1037
    first = arch.receptive_field - 1 - arch.stack_padding + 7 * arch.stack_stride  # frame index of the first logits
    T = first + 1 + 2 * arch.stack_stride  # three emissions
    gen = torch.Generator().manual_seed(100 + (1 if rnd else 0))
    x = torch.rand((N, 3, T, V, 2), generator=gen)
    pick = torch.randperm(N, generator=gen)[:60].tolist() + [0, 1, N - 2, N - 1]
    pick = sorted(set(pick))
    while len(pick) < 64:
        pick = sorted(set(pick + [int(torch.randint(0, N, (1,), generator=gen))]))
    xs = x[pick]
    ref = step.StepModel(sd, arch)
    with torch.no_grad():
        want = ref.forward_steps(xs)
        feats = []
        regular.stack_features(regular.normalise_input(xs, sd), sd, arch, feats, per_frame=per_frame)
    xd = x.to(DEV)
    outs = []
    for t in range(T):
        o = m.forward_step(xd[:, :, t].contiguous())
        assert m.last_schedule() == ref.trace[t], (tag, t)
        if o is not None:
            outs.append(o[pick].cpu())
    assert m.device_error() == 0, hex(m.device_error())
    assert m.tensor_core_blocks()[1:] == [3] * 9
    got = torch.stack(outs, dim=2)
    assert tuple(got.shape) == tuple(want.shape) == (len(pick), arch.classes, 3)
    err, mag = float((got - want).abs().max()), float(want.abs().max())
    print(f"\n[parity] {tag} {'randomised' if rnd else 'reference-init'} N={N} distinct streams, {len(pick)} checked: "
          f"max|d|={err:.3e} max|logit|={mag:.2f}")
    assert err <= 1e-3 * (max(1.0, mag / 16.0) if rnd else 1.0), (tag, rnd, err, mag)
    assert torch.equal(got.argmax(1), want.argmax(1))
    # block outputs of the picked streams (rows n*2 + person) at the last frame
    rows = torch.tensor([2 * n + s for n in pick for s in range(2)])
    for i in (4, 7, 9):
        n_out = sum(1 for f in ref.trace if f[i])
        blk = m.read_block(i)[rows.to(DEV)].cpu()
        e = _rel_err(blk, feats[i][:, :, n_out - 1])
        assert e < 2e-4, (tag, rnd, i, e)


@pytest.mark.parametrize("tag,rnd", [("cost_gcn", False), ("cost_gcn", True), ("cost_gcn_mod", True)])
def test_late_emissions_vs_oracle_nonconstant_input(tag, rnd):
    """Pooled logits long after the sliding window wrapped (> 2 x pool_size emissions) on a NON-constant input, against
    the oracle's exact window sum: a wrong ring slot or a drifting running sum shows up here."""
    arch, sd, m, V, _ = _build(tag, rnd, pool_size=20, pool_padding=5)
    n_emit = 2 * 20 + 7
    first = arch.receptive_field - 1 - arch.stack_padding + (20 - 1 - 5) * arch.stack_stride
    T = first + 1 + (n_emit - 1) * arch.stack_stride
    x = weights.make_input((2, 3, T, V, 2), seed=51)
    ref = step.StepModel(sd, arch)
    with torch.no_grad():
        want = ref.forward_steps(x)
    got = m.forward_steps(x.to(DEV))
    assert m.device_error() == 0
    assert tuple(got.shape) == tuple(want.shape) == (2, arch.classes, n_emit)
    got = got.cpu()
    err = (got - want).abs().amax(dim=(0, 1))
    mag = float(want.abs().max())
    print(f"\n[parity] {tag} late emissions: max|d| first={float(err[0]):.2e} after-wrap(max of last 20)={float(err[-20:].max()):.2e} "
          f"max|logit|={mag:.2f}")
    assert float(err.max()) <= 1e-3 * (max(1.0, mag / 16.0) if rnd else 1.0)
    assert torch.equal(got.argmax(1), want.argmax(1))


def test_default_pool_window_late_emissions():
    """The default CoST-GCN window (75 entries, 19 of padding): emissions 1, 57 (window exactly full of real entries)
    and 160 (> 2 windows later) against the oracle, single stream."""
    arch, sd, m, V, _ = _build("cost_gcn", False)
    n_emit = 160
    T = 297 + (n_emit - 1) * 4
    x = weights.make_input((1, 3, T, V, 2), seed=52)
    ref = step.StepModel(sd, arch)
    with torch.no_grad():
        want = ref.forward_steps(x)
    got = m.forward_steps(x.to(DEV)).cpu()
    assert m.device_error() == 0
    assert tuple(got.shape) == tuple(want.shape) == (1, 60, n_emit)
    for j in (0, 56, 57, 75, 151, n_emit - 1):
        assert float((got[:, :, j] - want[:, :, j]).abs().max()) <= 1e-3, j
    assert torch.equal(got.argmax(1), want.argmax(1))


@pytest.mark.parametrize("rnd", [False, True])
def test_coa_gcn_ntu120_vs_step_oracle(rnd):
    """BASELINE configs[2]: CoA-GCN on NTU RGB+D 120 (120 classes): the north-star clip against the step oracle."""
    arch, sd, m, V, _ = _build("coa_gcn", rnd, classes=120, dataset_name="ntu120")
    assert m.output_shape == (120,)
    x = weights.make_input((2, 3, 300, V, 2), seed=11)
    with torch.no_grad():
        want = step.StepModel(sd, arch).forward_steps(x)
    got = m.forward_steps(x.to(DEV)).cpu()
    assert m.device_error() == 0
    assert tuple(got.shape) == tuple(want.shape) == (2, 120)
    err, mag = float((got - want).abs().max()), float(want.abs().max())
    print(f"\n[parity] coa_gcn ntu120 {'randomised' if rnd else 'reference-init'}: max|d|={err:.3e} max|logit|={mag:.2f}")
    assert err <= 1e-3 * (max(1.0, mag / 16.0) if rnd else 1.0)
    assert torch.equal(got.argmax(1), want.argmax(1))


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two CUDA devices")
def test_sharded_logits_bit_identical_to_single_gpu():
    """streams sharded over 2 ranks (one process per GPU, NCCL all-gather of the logits) give the
    single-GPU logits bit for bit, with an odd stream count (uneven shards, different tile boundaries per rank)."""
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29613", os.path.join(ROOT, "tests", "dist_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "SHARDED_EQUALS_SINGLE ok" in r.stdout, r.stdout[-2000:]


# ---------------------------------------------------------------------------------------------
# time-batched forward_steps: module by module over chunks of frames
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag", ["cost_gcn", "cost_gcn_mod"])
def test_time_batched_forward_steps_bit_identical_and_few_launches(tag, monkeypatch):
    """forward_steps over a 300-frame clip with the stack walked module by module -- one launch per kernel and chunk, work
    items (frame, tile) -- gives bit for bit what stepping frame by frame gives (same kernels, same per-tile arithmetic),
    for the whole clip in one chunk, for a chunk size that makes the rings wrap inside the clip, and when stepping
    continues frame by frame afterwards; the whole clip takes at most 40 launches (models/base.py:187-190).
    The time-batched walk runs every block as graph conv + temporal conv launches (a launch covering many frames cannot
    feed frame n's temporal conv from the same launch's graph conv of frame n-1), so the bit-for-bit reference is the
    frame-by-frame path with the one-kernel block step switched off; against the default path the logits agree to rounding."""
    monkeypatch.setenv("COSK_FUSE_BLOCK", "0")
    cls, arch_fn, dataset, V, _ = MODELS[tag]
    arch = arch_fn()
    sd = weights.make_state_dict(arch, seed=8, randomize=True)
    x = weights.make_input((7, 3, 304, V, 2), seed=61).to(DEV)  # 14 skeletons: 3 tiles, the last one ragged

    def build(tc):
        m = cls({"dataset_name": dataset, "time_chunk": tc})
        m.load_state_dict(m.map_state_dict(sd), strict=True)
        return m

    ref = build(1)
    want = ref.forward_steps(x)
    per_frame_launches = ref.launch_count()
    assert ref.device_error() == 0 and want is not None and want.dim() == 3
    whole = build(-1)
    got = whole.forward_steps(x)
    assert whole.device_error() == 0
    assert whole._engine.time_chunk == 304
    assert whole.launch_count() <= 40, whole.launch_count()
    assert per_frame_launches > 50 * whole.launch_count()
    assert torch.equal(got, want)
    chunked = build(37)
    a = chunked.forward_steps(x[:, :, :250].contiguous())  # 6 full chunks + 28 frames
    outs = [] if a is None else ([a] if a.dim() == 2 else [a[:, :, j] for j in range(a.shape[2])])
    for t in range(250, 304):  # ... then frame by frame on the same state
        o = chunked.forward_step(x[:, :, t].contiguous())
        if o is not None:
            outs.append(o)
    assert chunked.device_error() == 0
    assert len(outs) == want.shape[2]
    for j, o in enumerate(outs):
        assert torch.equal(o, want[:, :, j]), j
    # and the end-of-clip flush on top of a batched clip
    ref.clean_state()
    full_ref = ref.forward_steps(x, pad_end=True)
    full = build(64).forward_steps(x, pad_end=True)
    assert torch.equal(full, full_ref)
    # the default frame-by-frame path (64-channel blocks as one kernel per step): same logits to rounding
    monkeypatch.delenv("COSK_FUSE_BLOCK")
    fused = build(1)
    assert "block" in fused_knobs(fused, x)
    got_f = fused.forward_steps(x)
    mag = float(want.abs().max())
    assert float((got_f - want).abs().max()) <= 2e-4 * max(1.0, mag), (float((got_f - want).abs().max()), mag)
    assert torch.equal(got_f.argmax(1), want.argmax(1))


def fused_knobs(model, x):
    """Names of the per-block kernel entries after the engine exists (one throw-away step creates it)."""
    model.forward_step(x[:, :, 0].contiguous())
    model.clean_state()
    return {k for blk in model.knobs()["blocks"] for k in blk}


def test_time_batched_strided_stack_kinetics(monkeypatch):
    """Headless three-block stack with every residual kind and a stride-2 block on the 18-joint skeleton: chunked in time
    (ring wrap, firing phase crossing chunk borders) == frame by frame, bit for bit (one-kernel block step off, see above),
    and == the oracle's clip math."""
    monkeypatch.setenv("COSK_FUSE_BLOCK", "0")
    blocks = [weights.BlockSpec(64, 64, 1, False), weights.BlockSpec(64, 64, 1, True), weights.BlockSpec(64, 128, 2, True)]
    arch = weights.ArchSpec(blocks, padding=4, skeleton="kinetics", head=False, block_names=["0.", "1.", "2."])
    sd = weights.make_state_dict(arch, seed=31, randomize=True)
    x = weights.make_input((23, 64, 45, 18), seed=32)
    mapped = {}
    for k, v in sd.items():
        i, rest = k.split(".", 1)
        kind = blocks[int(i)].res_kind
        mapped[f"{i}.{rest}" if kind == 0 else (f"{i}.0.0.{rest}" if rest.startswith("residual") else f"{i}.0.1.{rest}")] = v
    outs = {}
    for tc in (1, 7, 45):
        st = cs.CoStack([cs.BlockSpec(b.cin, b.cout, b.stride, b.residual) for b in blocks], padding=4, skeleton="kinetics", time_chunk=tc)
        st.load_state_dict(mapped, strict=True)
        outs[tc] = st.forward_steps(x.to(DEV))
        assert st.device_error() == 0
    assert torch.equal(outs[7], outs[1]) and torch.equal(outs[45], outs[1])
    with torch.no_grad():
        target = regular.stack_features(x, sd, arch)
    n = outs[1].shape[2]
    assert _rel_err(outs[1].cpu(), target[:, :, :n]) < 2e-4
