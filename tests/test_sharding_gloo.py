"""world_size-2 gloo test of the N>1 host logic: contigs are partitioned over ranks, every rank produces
the polished bytes of its shard (here with the oracle standing in for the GPU worker — this test is about
the sharding and the gather, not the kernels), rank 0 gathers and must reproduce the single-process result."""
import ctypes as C
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _polish_with_oracle(E, shard, cfg):
    O = C.CDLL(os.path.join(ROOT, "oracle", "libnp_oracle.so"))
    O.np_oracle_run.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    cap = int(shard.total_bases * 2) + 4096
    out = np.zeros(cap, np.uint8)
    off = np.zeros(shard.n_contigs + 1, np.int64)
    assert O.np_oracle_run(C.addressof(shard.view), 1, C.cast(cfg, C.c_void_p), out.ctypes.data, cap, off.ctypes.data) == 0
    return out[:off[-1]].copy(), off


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nextpolish_b200 import engine as E
    from nextpolish_b200.sharding import gather_bytes, partition_contigs
    kw = dict(seed=77, n_contigs=7, contig_len=0, min_len=500, max_len=9000, depth=15.0)
    full = E.Shard.synthetic(E.synth_params(**kw), 0, 7)
    lengths = [int(full.view.ctg_off[i + 1] - full.view.ctg_off[i]) for i in range(7)]
    mine = partition_contigs(lengths, world)[rank]
    cfg = E.default_config(b"")
    pieces = []
    for c in mine:                                    # contiguous [c, c+1) shards of the synthetic genome
        sh = E.Shard.synthetic(E.synth_params(**kw), c, c + 1)
        seq, _ = _polish_with_oracle(E, sh, cfg)
        pieces.append(seq)
    local = torch.from_numpy(np.concatenate(pieces) if pieces else np.zeros(0, np.uint8))
    got = gather_bytes(local, dst=0)
    from nextpolish_b200.sharding import FixedGather
    fg = FixedGather(200000, torch.device("cpu"), slots=2)
    buf = fg.send_buffer(torch.device("cpu"))
    buf[fg.HEADER:fg.HEADER + local.numel()] = local
    fg.set_count(buf, local.numel())
    fg(buf, slot=1)
    got2 = fg.result()
    if rank == 0:
        assert all(bytes(a.numpy()) == bytes(b.numpy()) for a, b in zip(got, got2))
        whole, off = _polish_with_oracle(E, full, cfg)
        parts = partition_contigs(lengths, world)
        ok = True
        for r in range(world):
            want = np.concatenate([whole[off[c]:off[c + 1]] for c in parts[r]]) if parts[r] else np.zeros(0, np.uint8)
            ok = ok and bytes(got[r].numpy()) == bytes(want)
        q.put(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_partition_is_balanced_and_complete():
    from nextpolish_b200.sharding import partition_contigs
    lengths = [1000000] * 5 + [20000, 30000, 250000, 777, 5]
    for n in (1, 2, 4, 8):
        parts = partition_contigs(lengths, n)
        assert sorted(i for p in parts for i in p) == list(range(len(lengths)))
        loads = [sum(lengths[i] for i in p) for p in parts]
        assert max(loads) - min(loads) <= max(lengths)


def test_two_rank_gloo_gather_reproduces_single_process(E, oracle):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=180)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert ok
