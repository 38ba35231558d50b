"""np_resident (resident_slots.cu): several engines polishing shards that already sit in HBM at the same time;
every job's bytes must equal what one engine delivers for the same shard and task."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _device_view(E, torch, sh, with_qual):
    a = sh.arrays()
    keep = {k: torch.from_numpy(v.copy()).cuda() for k, v in a.items() if k in ("ctg_seq", "rec_off", "rec", "qual_off", "qual")}
    v = E.ShardView()
    v.n_contigs, v.n_reads = sh.view.n_contigs, sh.view.n_reads
    v.ctg_off, v.ctg_read_off = sh.view.ctg_off, sh.view.ctg_read_off
    v.ctg_seq, v.rec_off, v.rec = keep["ctg_seq"].data_ptr(), keep["rec_off"].data_ptr(), keep["rec"].data_ptr()
    if with_qual:
        v.qual_off, v.qual = keep["qual_off"].data_ptr(), keep["qual"].data_ptr()
    return v, keep


@pytest.mark.parametrize("slots", [1, 3])
def test_resident_slots_equal_one_engine(E, slots):
    import torch
    cfg = E.default_config(b"")
    cfg.contents.read_tlen = 1750
    eng = E.Engine(0)
    shards = [E.Shard.synthetic(E.synth_params(seed=700 + k, n_contigs=2 + k, contig_len=120000, depth=25.0,
                                               lowercase_frac=0.002 * (k + 1)), 0, 2 + k, with_qual=2) for k in range(3)]
    views = [_device_view(E, torch, sh, True) for sh in shards]
    jobs = [(t, k) for k in range(len(shards)) for t in E.TASKS] * 3
    want = {}
    for t, k in set(jobs):
        got = eng.polish(shards[k], t, cfg)
        want[(t, k)] = b"".join(got[n] for n in shards[k].names)
    eng.close()
    cap = int(max(sh.total_bases for sh in shards) * 2) + 4096
    rp = E.ResidentSlots(0, slots)
    bufs = [torch.zeros(cap + 16, dtype=torch.uint8, device="cuda") for _ in range(slots)]
    pending = []

    def collect():
        tk, t, k = pending.pop(0)
        n = rp.wait(tk)
        b = bufs[tk % slots]
        assert int(b[:8].cpu().view(torch.int64).item()) == n == len(want[(t, k)])
        assert b[16:16 + n].cpu().numpy().tobytes() == want[(t, k)], (t, k, slots)

    for i, (t, k) in enumerate(jobs):            # tickets are consecutive from 0: job i lands in slot i % slots
        while len(pending) >= slots:
            collect()
        tk = rp.submit(t, views[k][0], cfg, bufs[i % slots].data_ptr(), cap + 16)
        assert tk == i
        pending.append((tk, t, k))
    while pending:
        collect()
    assert rp.launch_count() > 0
    with pytest.raises(E.NativeError):
        rp.wait(10 ** 6)
    rp.close()
