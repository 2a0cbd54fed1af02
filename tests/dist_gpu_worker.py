"""Worker of tests/test_gpu_scale.py::test_sharded_logits_bit_identical_to_single_gpu (launched with torchrun, one
process per GPU): every rank steps its shard of a seeded batch of distinct streams, logits are all-gathered over
NCCL, and rank 0 compares them bit for bit with stepping the whole batch on its own GPU."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import continual_skeletons_b200 as cs  # noqa: E402
from oracle import weights  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    arch = weights.cost_gcn_arch(pool_size=4, pool_padding=0)
    sd = weights.make_state_dict(arch, seed=8, randomize=True)
    N, T = 333, 76 + 3 * 4 + 1 + 8  # odd stream count: uneven shards; 3 emissions
    x = torch.rand((N, 3, T, 25, 2), generator=torch.Generator().manual_seed(5))

    def run(xs):
        m = cs.CoStGcn({"dataset_name": "dummy_ntu", "pool_size": 4, "pool_padding": 0})
        m.load_state_dict(m.map_state_dict(sd), strict=True)
        out = m.forward_steps(xs.to(dev))
        assert m.device_error() == 0
        return out

    lo, hi = cs.shard_range(N, rank, world)
    local_out = run(x[lo:hi])  # (n_local, classes, n_emissions)
    full = cs.all_gather_logits(local_out.contiguous(), N)
    ok = True
    if rank == 0:
        single = run(x)
        ok = tuple(full.shape) == tuple(single.shape) and bool(torch.equal(full, single))
        print(f"SHARDED_EQUALS_SINGLE {'ok' if ok else 'MISMATCH'} shape={tuple(full.shape)} world={world}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
