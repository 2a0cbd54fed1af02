"""Oracle restatement of the CONTINUAL per-step forward (the hot path), eager torch on CPU.

Test infrastructure only; also the CPU baseline that bench.py times (kind "port").

The continual semantics live in the un-vendored ``continual-inference`` package (>=0.16.0,
reference requirements.txt:3).  What is restated here is its published per-step algorithm as used
at the reference's call sites:

* ``co.Conv2d((9,1), padding=(p,0), stride=(s,1)).forward_step``  (models/base.py:318-324): keeps the
  last 8 inputs (zero initialised), emits at local step n iff ``n >= 8-p`` and ``(n-(8-p)) % s == 0``.
* ``co.Delay(d)`` (models/base.py:438) / the delay inside ``co.Residual`` (:417): FIFO that returns
  the value pushed d pushes earlier and nothing for its first d pushes.
* ``co.Residual(seq, residual_shrink=True)`` (:416-420): delay = seq.delay (8-p), halved when the
  sequence shrinks the clip (p == 0).
* ``co.BroadcastReduce(residual+align, gcn+tcn)`` (:432-444): sum, emitted when both branches emit.
* ``co.Sequential.forward_step``: a child that yields nothing short-circuits the rest.
* ``co.AvgPool1d(P, stride=1, padding=pp)`` (:97): zero-initialised window of P, first emission after
  ``P-1-pp`` earlier inputs; ``co.Linear`` (:99).

It is pinned against oracle/regular.py by the relations of the reference's tests
(tests/test_cost_gcn.py:37-362, tests/test_st_gcn_mod.py:11-90) in tests/test_oracle.py.
"""
from collections import deque

import torch
import torch.nn.functional as F

from . import regular

KT = 9  # temporal kernel size used everywhere in the reference (models/base.py:284,310)


class _Fifo:
    """co.Delay: output the value pushed ``depth`` pushes ago."""

    def __init__(self, depth):
        self.depth, self.q = depth, deque()

    def push(self, x):
        self.q.append(x)
        if len(self.q) > self.depth:
            return self.q.popleft()
        return None


class _StepConv:
    """co.Conv2d((k,1)) + BatchNorm2d, one frame at a time (CoTemporalConvolution, base.py:307-334)."""

    def __init__(self, sd, key, k, stride, pad):
        self.sd, self.key, self.k, self.stride, self.pad = sd, key, k, stride, pad
        self.window = None  # the last k-1 inputs
        self.n = 0

    def step(self, x):  # x (B, C, V)
        if self.window is None:
            self.window = [torch.zeros_like(x) for _ in range(self.k - 1)]
        frames = self.window + [x]
        first = self.k - 1 - self.pad
        fire = self.n >= first and (self.n - first) % self.stride == 0
        y = None
        if fire:
            clip = torch.stack(frames, dim=2)  # (B, C, k, V)
            w = self.sd[self.key + "t_conv.weight"].to(x.dtype)
            b = self.sd[self.key + "t_conv.bias"].to(x.dtype)
            y = F.conv2d(clip, w, b)  # valid conv over exactly k frames -> T = 1
            y = regular._bn(y, self.sd, self.key + "bn.")[:, :, 0]
        self.window = frames[1:]
        self.n += 1
        return y


class StepBlock:
    """One CoSpatioTemporalBlock (models/base.py:390-446) advanced frame by frame."""

    def __init__(self, sd, key, spec, pad):
        self.sd, self.key, self.spec = sd, key, spec
        self.tcn = _StepConv(sd, key + "tcn.", KT, spec.stride, pad)
        shrink = (KT - 2 * pad) != 1
        tcn_delay = KT - 1 - pad
        if spec.res_kind == 1:
            d = tcn_delay // 2 if shrink else tcn_delay
            self.align = _Fifo(d)
        elif spec.res_kind == 2:
            d = tcn_delay // spec.stride
            if shrink:
                d //= 2
            self.res = _StepConv(sd, key + "residual.", 1, spec.stride, 0)
            self.align = _Fifo(d)

    def step(self, x):  # (B, Cin, V) -> (B, Cout, V) | None
        g = regular.graph_conv(x.unsqueeze(2), self.sd, self.key + "gcn.")[:, :, 0]
        z = self.tcn.step(g)
        kind = self.spec.res_kind
        if kind == 0:
            out = z
        else:
            if kind == 1:
                r = self.align.push(x)
            else:
                r = self.res.step(x)
                if r is not None:
                    r = self.align.push(r)
            out = (z + r) if (z is not None and r is not None) else None
        return None if out is None else F.relu(out)


class StepModel:
    """CoStGcn / CoStGcnMod ``forward_step`` / ``forward_steps`` (models/base.py:183-190) on CPU."""

    def __init__(self, sd, arch):
        self.sd, self.arch = sd, arch
        self.clean_state()

    def clean_state(self):
        a = self.arch
        self.blocks = [StepBlock(self.sd, n, s, a.padding) for n, s in zip(a.block_names, a.blocks)]
        self.pool_window = None
        self.pool_n = 0
        self.frame = 0
        self.trace = []  # per frame: tuple of per-block emit flags + head flag (schedule parity)

    def forward_step(self, x):
        """x (N, C, V, M) -> (N, classes) | None.  With ``arch.head == False`` the input is
        (B, C, V) and the last block's output is returned."""
        a, sd = self.arch, self.sd
        if a.head:
            N, C, V, M = x.shape
            f = x.permute(0, 3, 2, 1).contiguous().view(N, M * V * C)  # base.py:73-75
            f = regular._bn(f, sd, "data_bn.")
            h = f.view(N, M, V, C).permute(0, 1, 3, 2).contiguous().view(N * M, C, V)  # :77-82
        else:
            h = x
        flags = []
        for blk in self.blocks:
            h = blk.step(h) if h is not None else None
            flags.append(h is not None)
        out = None
        if a.head and h is not None:
            N = x.shape[0]
            c = h.shape[1]
            p = h.view(N, a.persons, c, a.vertices).mean(3).mean(1)  # base.py:84
            if self.pool_window is None:
                self.pool_window = deque(torch.zeros_like(p) for _ in range(a.pool_size - 1))
            self.pool_window.append(p)
            if self.pool_n >= a.pool_size - 1 - a.pool_padding:
                mean = torch.stack(list(self.pool_window), 0).sum(0) / a.pool_size
                out = F.linear(mean, sd["fc.weight"].to(p.dtype), sd["fc.bias"].to(p.dtype))
            self.pool_window.popleft()
            self.pool_n += 1
        elif not a.head:
            out = h
        self.trace.append(tuple(flags) + (out is not None,))
        self.frame += 1
        return out

    def forward_steps(self, x, pad_end=False):
        """x (N, C, T, V, M) [or (B, C, T, V) without head] -> stacked emissions along a new last
        axis, squeezed when there is exactly one and the stack has a head (base.py:101)."""
        outs = []
        T = x.shape[2]
        for t in range(T):
            o = self.forward_step(x[:, :, t])
            if o is not None:
                outs.append(o)
        if pad_end:
            # The library flushes every temporal conv with its own end padding, module by module;
            # only the reference's block-level tests use it (never CoModelBase.forward,
            # models/base.py:177), so it is outside the restated path.
            raise NotImplementedError("pad_end=True is not part of the restated hot path")
        if not outs:
            return None
        y = torch.stack(outs, dim=2)
        if self.arch.head and y.shape[2] == 1:
            y = y[:, :, 0]
        return y
