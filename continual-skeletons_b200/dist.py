"""Multi-GPU plumbing: streams are independent, so the batch of streams is sharded across ranks
with no traffic on the step; the only collective is an all-gather of logits on emitting steps
(SURVEY.md section 8e).  One process per GPU, ``torch.distributed`` (NCCL on GPUs, gloo in the CPU
tests)."""
import torch
import torch.distributed as dist


def shard_range(n_streams, rank, world):
    """Contiguous, balanced split: the first ``n % world`` ranks hold one extra stream."""
    base, extra = divmod(int(n_streams), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def any_rank(flag, device="cpu", group=None):
    """True on every rank iff ``flag`` is true on at least one.  For loops whose length is decided by something rank-local
    (a wall clock, a queue) while their body carries a collective: all ranks must run the same number of iterations, or one
    of them leaves the loop early and meets its peers' all-gather with a different collective (which deadlocks NCCL)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return bool(flag)
    t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return bool(int(t.item()))


def all_gather_logits(local, n_streams, group=None):
    """local (n_local, classes) -> (n_streams, classes) on every rank, in global stream order."""
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world = dist.get_world_size(group)
    if world == 1:
        return local
    counts = [hi - lo for lo, hi in (shard_range(n_streams, r, world) for r in range(world))]
    width = max(counts)
    tail = tuple(local.shape[1:])
    send = local.contiguous()
    if send.shape[0] < width:  # uneven split: pad to the widest shard so every backend accepts it
        send = torch.cat([send, send.new_zeros((width - send.shape[0],) + tail)], 0)
    bufs = [torch.empty((width,) + tail, dtype=local.dtype, device=local.device) for _ in range(world)]
    dist.all_gather(bufs, send, group=group)
    return torch.cat([b[:c] for b, c in zip(bufs, counts)], 0)


class LogitGather:
    """The same all-gather for a serving loop: buffers allocated once, the collective issued on a side stream so
    that the next step's kernels do not wait for it ("overlap with the next step on a side
    stream").  ``launch(local)`` returns immediately; ``result()`` makes the current stream wait for the newest
    gather and returns the (n_streams, classes) matrix in global stream order.  Two result buffers alternate, so a
    gather may still be in flight while the previous result is being read."""

    def __init__(self, n_streams, classes, device, dtype=torch.float32, group=None):
        self.group = group
        self.active = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self.n_streams, self.classes = int(n_streams), int(classes)
        self.device = torch.device(device)
        self._last = None
        if not self.active:
            return
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        self.counts = [hi - lo for lo, hi in (shard_range(n_streams, r, world) for r in range(world))]
        self.width = max(self.counts)
        self.even = all(c == self.width for c in self.counts)
        self.n_local = self.counts[rank]
        self.stream = torch.cuda.Stream(device=self.device) if self.device.type == "cuda" else None
        self.send = torch.zeros((self.width, classes), dtype=dtype, device=self.device)
        self.recv = [torch.empty((world * self.width, classes), dtype=dtype, device=self.device) for _ in range(2)]
        self.out = [torch.empty((n_streams, classes), dtype=dtype, device=self.device) for _ in range(2)]
        self.done = [torch.cuda.Event() for _ in range(2)] if self.stream is not None else [None, None]
        self.turn = 0

    def launch(self, local):
        if not self.active:
            self._last = local
            return
        b = self.turn
        self.turn ^= 1
        if self.stream is None:  # CPU / gloo: synchronous
            self.send[: self.n_local].copy_(local)
            dist.all_gather_into_tensor(self.recv[b], self.send, group=self.group)
            self._compact(b)
            self._last = b
            return
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ready)  # the step's logits are complete
            local.record_stream(self.stream)
            self.send[: self.n_local].copy_(local, non_blocking=True)
            dist.all_gather_into_tensor(self.recv[b], self.send, group=self.group)
            self._compact(b)
            self.done[b].record(self.stream)
        self._last = b

    def _compact(self, b):
        if self.even:
            return
        lo = 0
        for r, c in enumerate(self.counts):  # drop the padding rows of the short shards
            self.out[b][lo:lo + c].copy_(self.recv[b][r * self.width:r * self.width + c], non_blocking=True)
            lo += c

    def result(self):
        """Gathered logits of the newest ``launch`` (the current stream is made to wait for them)."""
        if not self.active:
            return self._last
        b = self._last
        if b is None:
            return None
        if self.stream is not None:
            torch.cuda.current_stream(self.device).wait_event(self.done[b])
        return self.recv[b] if self.even else self.out[b]
