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
    """The same collective with buffers allocated once, ONE collective per call and no host synchronisation in the
    hot loop: every rank contributes a capacity-sized buffer whose first HEADER bytes hold its byte count
    (little-endian int64) and whose payload follows; rank `dst` slices afterwards.  `slots` rotating receive
    sets let a gather stay in flight while the next step is being computed.

    A true gather (dist.gather: grouped ncclSend / ncclRecv on GPUs, gloo in the CPU tests): only rank `dst` receives —
    N x (cap + HEADER) bytes per call — every other rank sends its own buffer once.  (Round 1 issued an all-gather, which
    delivered all N buffers to every rank.)"""
    HEADER = 16

    def __init__(self, cap, device, dst=0, group=None, slots=1):
        self.cap, self.dst, self.group = cap, dst, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        root = self.rank == dst
        n = cap + self.HEADER
        self.bufs = [[torch.empty(n, dtype=torch.uint8, device=device) for _ in range(self.world)]
                     for _ in range(slots)] if root else None
        self.last = 0

    def send_buffer(self, device):
        """A buffer of the right size for __call__ (header + cap payload bytes)."""
        return torch.zeros(self.cap + self.HEADER, dtype=torch.uint8, device=device)

    @staticmethod
    def set_count(buf, n):
        """Writes the byte count into the header (host-side helper for CPU tensors / small jobs)."""
        buf[:8] = torch.tensor([n], dtype=torch.int64).view(torch.uint8).to(buf.device)

    def __call__(self, buf, slot=0, async_op=False):
        """buf: uint8 tensor of cap + HEADER elements (header already filled)."""
        self.last = slot
        return dist.gather(buf, self.bufs[slot] if self.bufs else None, dst=self.dst, group=self.group, async_op=async_op)

    def result(self, slot=None):
        """On dst: list of per-rank byte tensors of the last call (synchronises)."""
        if self.rank != self.dst:
            return None
        out = []
        for b in self.bufs[self.last if slot is None else slot]:
            n = int(b[:8].cpu().view(torch.int64).item())
            out.append(b[self.HEADER:self.HEADER + n])
        return out
