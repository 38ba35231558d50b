"""Contig sharding across GPUs and the single collective of the path.

The reference shards contigs into blocks by cumulative length (source/nextPolish:93-117, consumed by
nextpolish1.py -b/-i); here the same unit (a whole contig) is assigned to ranks with a longest-first
greedy so that bases per rank are balanced.  After a task step every rank holds the polished bytes of
its contigs; `gather_bytes` moves them to rank 0 with one size exchange + one gather (NCCL on GPUs,
gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def partition_contigs(lengths, n_shards):
    """Longest-processing-time greedy: returns n_shards lists of contig indices (each sorted)."""
    order = sorted(range(len(lengths)), key=lambda i: (-lengths[i], i))
    loads = [0] * n_shards
    shards = [[] for _ in range(n_shards)]
    for i in order:
        k = min(range(n_shards), key=lambda j: (loads[j], j))
        shards[k].append(i)
        loads[k] += lengths[i]
    return [sorted(s) for s in shards]


def gather_bytes(local, dst=0, group=None):
    """Gather variable-length uint8 tensors to `dst`. Returns the list (by rank) on dst, None elsewhere."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n = torch.tensor([local.numel()], dtype=torch.int64, device=local.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    cap = max(max(sizes), 1)
    buf = torch.zeros(cap, dtype=torch.uint8, device=local.device)
    buf[:local.numel()] = local
    out = [torch.empty(cap, dtype=torch.uint8, device=local.device) for _ in range(world)] if rank == dst else None
    dist.gather(buf, out, dst=dst, group=group)
    if rank != dst:
        return None
    return [o[:s] for o, s in zip(out, sizes)]


class FixedGather:
    """The same collective with buffers allocated once and no host synchronisation in the hot loop:
    every rank contributes a capacity-sized buffer plus its byte count; rank `dst` slices afterwards."""

    def __init__(self, cap, device, dst=0, group=None):
        self.cap, self.dst, self.group = cap, dst, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.n = torch.zeros(1, dtype=torch.int64, device=device)
        root = self.rank == dst
        self.sizes = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(self.world)] if root else None
        self.bufs = [torch.empty(cap, dtype=torch.uint8, device=device) for _ in range(self.world)] if root else None

    def __call__(self, buf, n):
        """buf: uint8 tensor of exactly `cap` elements holding n valid bytes."""
        self.n.fill_(n)
        dist.gather(self.n, self.sizes, dst=self.dst, group=self.group)
        dist.gather(buf, self.bufs, dst=self.dst, group=self.group)

    def result(self):
        """On dst: list of per-rank byte tensors of the last call (synchronises)."""
        if self.rank != self.dst:
            return None
        return [b[:int(s.item())] for b, s in zip(self.bufs, self.sizes)]
