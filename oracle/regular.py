"""Oracle restatement of the reference's NON-continual (clip) ST-GCN math, functional form.

Test infrastructure only.  Operates directly on a state_dict in the regular ``StGcn`` key format
(oracle/weights.py), eval mode (running BN statistics), torch CPU, dtype of the inputs (fp32 for
parity, fp64 for the noise floor).

Follows: models/base.py:260-270 (GraphConvolution.forward), :302-304 (TemporalConvolution.forward),
:376-387 (SpatioTemporalBlock.forward), models/st_gcn/st_gcn.py:48-65 (StGcn.forward) and
models/st_gcn_mod/st_gcn_mod.py:52-69 (StGcnMod.forward, identical body), models/base.py:73-101
(CoModelBase clip-mode pipeline: data_bn, blocks, spatial mean, AvgPool1d, fc).
"""
import torch
import torch.nn.functional as F

BN_EPS = 1e-5  # torch.nn.BatchNorm default, never overridden by the reference


def _bn(x, sd, key):
    w, b = sd[key + "weight"].to(x.dtype), sd[key + "bias"].to(x.dtype)
    m, v = sd[key + "running_mean"].to(x.dtype), sd[key + "running_var"].to(x.dtype)
    return F.batch_norm(x, m, v, w, b, training=False, eps=BN_EPS)


def _conv(x, sd, key, stride=1, pad=0):
    return F.conv2d(x, sd[key + "weight"].to(x.dtype), sd[key + "bias"].to(x.dtype), stride=(stride, 1), padding=(pad, 0))


def adaptive_graph_conv(x, sd, key):
    """AdaptiveGraphConvolution.forward, models/a_gcn/a_gcn.py:48-69: learned dense ``A + graph_attn`` plus a
    per-sample vertex attention softmax(theta^T phi / (inter_c*T)) over the source vertex (dim -2).  The
    attention spans all T frames of ``x``; the continual model (models/coa_gcn/coa_gcn.py:11-14,
    ``co.forward_stepping``) calls it with T = 1."""
    B, C, T, V = x.shape
    adj = (sd[key + "A"] + sd[key + "graph_attn"]).to(x.dtype)
    total = None
    for part in range(3):
        theta = _conv(x, sd, key + f"a_conv.{part}.")  # (B, inter_c, T, V)
        phi = _conv(x, sd, key + f"b_conv.{part}.")
        inter_c = theta.shape[1]
        a1 = theta.permute(0, 3, 1, 2).reshape(B, V, inter_c * T)
        a2 = phi.reshape(B, inter_c * T, V)
        att = torch.softmax(torch.matmul(a1, a2) / (inter_c * T), dim=-2) + adj[part]
        mixed = torch.matmul(x.reshape(B, C * T, V), att).reshape(B, C, T, V)
        z = _conv(mixed, sd, key + f"g_conv.{part}.")
        total = z + total if total is not None else z
    total = _bn(total, sd, key + "bn.")
    if (key + "gcn_residual.0.weight") in sd:
        skip = _bn(_conv(x, sd, key + "gcn_residual.0."), sd, key + "gcn_residual.1.")
    else:
        skip = x
    return F.relu(total + skip)


ATTENTION_HEADS = 8  # Nh default of GcnUnitAttention, never overridden (models/s_tr/s_tr.py:311; cos_tr.py:25-28)


def attention_graph_conv(x, sd, key):
    """GcnUnitAttention.forward with only_attention=True, eval mode (models/s_tr/s_tr.py:417-476) around
    SpatialAttention.forward (:134-231; relative=False, adjacency=False, drop-connect is training only):
    per frame, every vertex attends over the V vertices of its skeleton with 8 heads."""
    B, C, T, V = x.shape
    y = x.permute(0, 1, 3, 2).reshape(B, C * V, T)
    y = _bn(y, sd, key + "data_bn.")
    y = y.reshape(B, C, V, T).permute(0, 1, 3, 2)
    xa = y.permute(0, 2, 1, 3).reshape(B * T, C, 1, V)
    qkv = _conv(xa, sd, key + "attention_conv.qkv_conv.")
    dv = sd[key + "attention_conv.attn_out.weight"].shape[0]
    dk = (qkv.shape[1] - dv) // 2
    nh = ATTENTION_HEADS
    q, k, v = torch.split(qkv, [dk, dk, dv], dim=1)
    q = q.reshape(B * T, nh, dk // nh, V) * ((dk // nh) ** -0.5)
    k = k.reshape(B * T, nh, dk // nh, V)
    v = v.reshape(B * T, nh, dv // nh, V)
    w = torch.softmax(torch.matmul(q.transpose(2, 3), k), dim=-1)  # (BT, nh, V, V): row i attends over j
    o = torch.matmul(w, v.transpose(2, 3))  # (BT, nh, V, dvh)
    o = o.reshape(B * T, nh, 1, V, dv // nh).permute(0, 1, 4, 2, 3).reshape(B * T, dv, 1, V)
    o = _conv(o, sd, key + "attention_conv.attn_out.")
    o = o.reshape(B, T, dv, V).permute(0, 2, 1, 3)
    if C == dv:  # skip_conn and in_channels == out_channels
        o = o + x
    return F.relu(_bn(o, sd, key + "bn."))


def graph_conv(x, sd, key, per_frame=False):
    """x (B, Cin, T, V) -> (B, Cout, T, V).  models/base.py:260-270; dispatches to the adaptive variant when
    the state_dict holds its embedding convs.  ``per_frame`` evaluates the adaptive attention one frame at a
    time, which is what the continual model does step by step."""
    if (key + "attention_conv.qkv_conv.weight") in sd:
        return attention_graph_conv(x, sd, key)
    if (key + "a_conv.0.weight") in sd:
        if per_frame and x.shape[2] > 1:
            return torch.cat([adaptive_graph_conv(x[:, :, t: t + 1], sd, key) for t in range(x.shape[2])], dim=2)
        return adaptive_graph_conv(x, sd, key)
    B, C, T, V = x.shape
    adj = (sd[key + "A"] * sd[key + "graph_attn"]).to(x.dtype)
    total = None
    for part in range(3):
        mixed = torch.matmul(x.reshape(B, C * T, V), adj[part]).reshape(B, C, T, V)
        z = _conv(mixed, sd, key + f"g_conv.{part}.")
        total = z if total is None else z + total
    total = _bn(total, sd, key + "bn.")
    if (key + "gcn_residual.0.weight") in sd:
        skip = _bn(_conv(x, sd, key + "gcn_residual.0."), sd, key + "gcn_residual.1.")
    else:
        skip = x
    return F.relu(total + skip)


def temporal_conv(x, sd, key, stride, pad):
    """models/base.py:291-304."""
    return _bn(_conv(x, sd, key + "t_conv.", stride, pad), sd, key + "bn.")


def st_block(x, sd, key, spec, pad, per_frame=False):
    """models/base.py:376-387; ``pad`` is temporal_padding (4 regular / 0 for the * variant)."""
    z = temporal_conv(graph_conv(x, sd, key + "gcn.", per_frame), sd, key + "tcn.", spec.stride, pad)
    shrink = 4 - pad
    xs = x[:, :, shrink: x.shape[2] - shrink] if shrink else x
    if spec.res_kind == 0:
        r = 0
    elif spec.res_kind == 1:
        r = xs
    else:
        r = temporal_conv(xs, sd, key + "residual.", spec.stride, 0)
    return F.relu(z + r)


def normalise_input(x, sd):
    """(N, C, T, V, M) -> data_bn -> (N*M, C, T, V).  st_gcn.py:49-57 / base.py:73-82."""
    N, C, T, V, M = x.shape
    y = x.permute(0, 4, 3, 1, 2).contiguous().view(N, M * V * C, T)
    y = _bn(y, sd, "data_bn.")
    return y.view(N, M, V, C, T).permute(0, 1, 3, 4, 2).contiguous().view(N * M, C, T, V)


def stack_features(x, sd, arch, collect=None, per_frame=False):
    """Run all blocks on a clip (N*M, C, T, V); optionally collect per-block outputs."""
    for name, spec in zip(arch.block_names, arch.blocks):
        x = st_block(x, sd, name, spec, arch.padding, per_frame)
        if collect is not None:
            collect.append(x)
    return x


def stgcn_forward(x, sd, arch, collect=None):
    """Regular StGcn / StGcnMod clip forward -> (N, classes).  st_gcn.py:48-65."""
    N, M = x.shape[0], x.shape[4]
    y = stack_features(normalise_input(x, sd), sd, arch, collect)
    c = y.shape[1]
    y = y.view(N, M, c, -1).mean(3).mean(1)
    return F.linear(y, sd["fc.weight"].to(y.dtype), sd["fc.bias"].to(y.dtype))


def pooled_sequence(x, sd, arch, per_frame=False):
    """(N, C, T, V, M) -> spatially pooled last-block features (N, Cl, T_out).  base.py:84."""
    N, M = x.shape[0], x.shape[4]
    y = stack_features(normalise_input(x, sd), sd, arch, per_frame=per_frame)
    _, c, t, v = y.shape
    return y.view(N, M, c, t, v).mean(4).mean(1)


def co_clip_forward(x, sd, arch, per_frame=False):
    """What CoStGcn/CoStGcnMod ``forward`` returns in clip mode: AvgPool1d(pool_size, 1,
    pool_padding) over the pooled sequence, fc per time step, first step kept
    (models/base.py:97-101,166-181)."""
    h = pooled_sequence(x, sd, arch, per_frame)
    h = F.avg_pool1d(h, arch.pool_size, stride=1, padding=arch.pool_padding)  # count_include_pad
    logits = torch.einsum("kc,nct->nkt", sd["fc.weight"].to(h.dtype), h) + sd["fc.bias"].to(h.dtype)[None, :, None]
    return logits[:, :, 0]
