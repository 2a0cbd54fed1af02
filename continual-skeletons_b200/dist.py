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
